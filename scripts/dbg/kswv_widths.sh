# developer check: the kswv GPU tests at the default width choice, then the bench at each minimum width
timeout 600 python -m pytest tests/test_kswv_gpu.py -x -q 2>&1 | tail -2
for w in 8 16 32; do KSWV_MIN_LANES=$w timeout 200 python scripts/kswv_bench.py --pairs 200000 --steps 5 --warmup 1 --cpu-sample 2000 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('min lanes $w', 'kernel %.0f e2e %.0f GCUPS roofline %.3f'%(d['value'], d['e2e']['value'], d['roofline']['frac']))"; done
