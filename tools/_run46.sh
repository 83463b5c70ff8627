cd /root/repo
timeout 600 python -m pytest tests/test_tc_linear_gpu.py -x -q --timeout 300 2>&1 | tail -3
timeout 300 python tools/bench_tc_instep.py 2>&1 | tail -9
timeout 300 python tools/la_probe.py 2>&1 | grep "flush"
