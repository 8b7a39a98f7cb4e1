#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
echo "== all gpu tests"
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== driver-like bench"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2m_bench_drv.json 2> gpurun_out/r2m_bench_drv.err; tail -c 3000 gpurun_out/r2m_bench_drv.json; tail -3 gpurun_out/r2m_bench_drv.err
for i in 1 2 3; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-configs | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('20/5', d['ms_per_step'], d['roofline']['frac'], d['clocks'])"; done
timeout 300 python bench.py --steps 2000 --warmup 20 --no-cpu --no-e2e --no-configs | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('2000/20', d['ms_per_step'], d['roofline']['frac'], d['clocks'])"
echo "== reference arm"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 | cut -c1-600
echo "== sanitizer"
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r2m_memcheck_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_memcheck_smoke.out 2>&1; echo "memcheck smoke rc=$?"; tail -2 gpurun_out/r2m_memcheck_smoke.log
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2m_racecheck_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_racecheck_smoke.out 2>&1; echo "racecheck smoke rc=$?"; tail -2 gpurun_out/r2m_racecheck_smoke.log
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2m_racecheck_autoreset.log python -m pytest tests/test_cuda_parity.py -q -x -m gpu -k "prepared_resets and 4v8" > gpurun_out/r2m_racecheck_autoreset.out 2>&1; echo "racecheck autoreset rc=$?"; tail -2 gpurun_out/r2m_racecheck_autoreset.log; tail -2 gpurun_out/r2m_racecheck_autoreset.out
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2m_memcheck_autoreset.log python -m pytest tests/test_cuda_parity.py -q -x -m gpu -k "prepared_resets and 4v8" > gpurun_out/r2m_memcheck_autoreset.out 2>&1; echo "memcheck autoreset rc=$?"; tail -2 gpurun_out/r2m_memcheck_autoreset.log; tail -2 gpurun_out/r2m_memcheck_autoreset.out
