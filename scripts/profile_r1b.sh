set -x
mkdir -p gpurun_out/prof
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dropin --no-optimum > gpurun_out/prof/b_logw.json 2>/dev/null && \
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof/launches_logw.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dropin --no-optimum > /dev/null 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stream_pass -s 4 -c 4 -o gpurun_out/prof/full_stream_pass -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dropin --no-optimum > /dev/null 2>&1
timeout 150 python bench.py --method forces --steps 3 --warmup 3 --no-cpu-baseline --no-dropin --no-optimum > gpurun_out/prof/b_forces.json 2>/dev/null && \
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof/launches_forces.csv python bench.py --method forces --steps 3 --warmup 3 --no-cpu-baseline --no-dropin --no-optimum > /dev/null 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_team -s 4 -c 4 -o gpurun_out/prof/full_fused_team -f python bench.py --method forces --steps 3 --warmup 3 --no-cpu-baseline --no-dropin --no-optimum > /dev/null 2>&1
ls -la gpurun_out/prof
