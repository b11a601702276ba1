#!/bin/bash
# round 2, call V2 (one GPU): bench lines of all five workloads + the reference arm + variants with the final library
mkdir -p gpurun_out
rm -f gpurun_out/r02v_bench_lines.jsonl gpurun_out/r02v_bench_variants.jsonl
for w in auni3d dendrite2d auni2d gg3d_hbsm pfhub1a; do
  timeout -k 5 400 python bench.py --workload $w >> gpurun_out/r02v_bench_lines.jsonl 2> gpurun_out/r02v_bench_$w.err
done
timeout -k 5 400 python bench.py --impl reference >> gpurun_out/r02v_bench_lines.jsonl 2>> gpurun_out/r02v_bench_auni3d.err
timeout -k 5 200 python bench.py --workload auni3d --fd-flag 1 --no-e2e --no-cpu-baseline >> gpurun_out/r02v_bench_variants.jsonl 2>> gpurun_out/r02v_bench_auni3d.err
timeout -k 5 200 python bench.py --workload auni3d --newton warm --no-e2e --no-cpu-baseline --no-extras >> gpurun_out/r02v_bench_variants.jsonl 2>> gpurun_out/r02v_bench_auni3d.err
python - <<PY
import json
for f in ('gpurun_out/r02v_bench_lines.jsonl','gpurun_out/r02v_bench_variants.jsonl'):
  for l in open(f):
    if l.startswith('{'):
        d=json.loads(l)
        if d.get('impl')=='reference': print('REFERENCE', d['value'], d['cpu_baseline']['sample']); continue
        r=d['roofline']
        print(d['config']['workload'][:60], '| ms %.4f GCUPS %.2f frac %.3f fp64 %s e2e %s cpu %s'%(d['ms_per_step'], d['value'], r['frac'], r.get('fp64') and round(r['fp64']['frac'],3), d['e2e'] and round(d['e2e']['value'],3), d['cpu_baseline'] and d['cpu_baseline']['value']))
PY
