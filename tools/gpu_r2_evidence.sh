cd $GRAFT_REPO_ROOT
# 1. launch list of the benchmarked command (cold-cache, serialised times: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --headline-only --no-cpu > gpurun_out/r2_launches.log 2>&1
tail -1 gpurun_out/r2_launches.log | cut -c1-200
# 2. every kernel of the hot-path step, full sections (last step of 4)
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"fk_table|deform_preprocess|tile_|composite_|preprocess_bwd|lbs_|preprocess_scan|duplicate_keys|assemble" -s 34 -c 10 -o gpurun_out/r2_step python tools/prof_step.py --steps 4 > gpurun_out/r2_step.log 2>&1
tail -2 gpurun_out/r2_step.log
# 3. the kernels around the path: joint MLP, loss, Adam, densify statistics (complete eager iteration)
timeout 900 ncu --set full --clock-control none -k regex:"adam_kernel|ssim_|joint_|small_gemm|densify_stats" -c 14 -o gpurun_out/r2_iter python tools/train_once.py > gpurun_out/r2_iter.log 2>&1
tail -2 gpurun_out/r2_iter.log
# 4. widening rows: sp-stage LBS, densification
timeout 900 ncu --set full --clock-control none -k regex:"sp_table|lbs_fwd|lbs_bwd_kernel|sp_bwd|densify_|opacity_reset" -s 6 -c 10 -o gpurun_out/r2_widen python tools/prof_widening.py > gpurun_out/r2_widen.log 2>&1
tail -2 gpurun_out/r2_widen.log
ls -la gpurun_out/*.ncu-rep
