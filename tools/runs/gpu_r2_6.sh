cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"composite|onesweep" -s 16 -c 8 -o gpurun_out/r2_prof2 python tools/prof_step.py > gpurun_out/r2_ncu2.log 2>&1
tail -2 gpurun_out/r2_ncu2.log
