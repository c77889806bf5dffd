mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -k "bit_matrix_in_global or fused_and_unfused or full_pipeline or many_frames" 2>&1 | tail -5
timeout 100 python tools/fused_timeline.py 840 2>&1 | grep "BFS\|task total\|O(it) staged\|drosa_orient" | tee gpurun_out/r2/fused_timeline_840_after.log
timeout 100 python tools/stream_stats.py 200 5 2>&1 | tail -6 | tee gpurun_out/r2/stream_stats_after.log
