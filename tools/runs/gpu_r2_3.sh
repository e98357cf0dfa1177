cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"composite|onesweep|preprocess_scan" -s 18 -c 9 -o gpurun_out/r2_prof1 python tools/prof_step.py > gpurun_out/r2_ncu1.log 2>&1
tail -3 gpurun_out/r2_ncu1.log
ls -la gpurun_out/*.ncu-rep | tail -2
