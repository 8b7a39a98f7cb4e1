#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
echo "== all gpu tests"
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== PDL A/B"
for pdl in 1 0 1 0; do
MATE_B200_PDL=$pdl timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-configs | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pdl=$pdl 20/5', d['ms_per_step'], d['roofline']['frac'])"
done
for pdl in 1 0; do
MATE_B200_PDL=$pdl timeout 300 python bench.py --steps 2000 --warmup 20 --no-cpu --no-e2e --no-configs | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pdl=$pdl 2000/20', d['ms_per_step'], d['roofline']['frac'], d['episode_stats'])"
done
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2n_racecheck_autoreset.log python -m pytest tests/test_cuda_parity.py -q -x -m gpu -k "prepared_resets and 4v8" > gpurun_out/r2n_racecheck_autoreset.out 2>&1; echo "racecheck autoreset rc=$?"; tail -2 gpurun_out/r2n_racecheck_autoreset.log; tail -2 gpurun_out/r2n_racecheck_autoreset.out
