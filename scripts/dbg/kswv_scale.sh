# developer run: kswv bench at N = 1, 2, 4 (.. 8) GPUs of one box, 400k pairs per GPU
for n in "$@"; do
  timeout 300 python scripts/kswv_bench.py --pairs $((400000*n)) --gpus $n --steps 10 --warmup 3 --cpu-sample $([ $n = 1 ] && echo 20000 || echo 500) > gpurun_out/kswv_n$n.json 2> gpurun_out/kswv_n$n.err || tail -5 gpurun_out/kswv_n$n.err
  python -c "
import json; d=json.load(open('gpurun_out/kswv_n$n.json')); print('N=$n value %.0f e2e %.0f frac %.3f parity %s host %s cpu %.0f'%(d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity']['mismatches'], d['run']['host_ms_last_step'], d['cpu_baseline']['value']))"
done
