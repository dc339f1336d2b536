python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "staged or ordered or constant or cs16" 2>&1 | tail -5
for st in 0 1; do AIRGPU_STAGE=$st python bench.py --no-cpu --no-e2e --no-extras --steps 10 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stage', $st, d['roofline']['kernel_ms'], d['roofline']['frac'], d['ms_per_step'], d['frames_per_step'])"; done
for st in 0 1; do AIRGPU_TRAFFIC=sparse AIRGPU_STAGE=$st python bench.py --no-cpu --no-e2e --no-extras --steps 10 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sparse stage', $st, d['roofline']['kernel_ms'], d['roofline']['frac'], d['ms_per_step'], d['frames_per_step'])"; done
python tools/measure_extras.py 2>&1 | tail -1
