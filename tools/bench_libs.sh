#!/bin/bash
# usage: tools/bench_libs.sh lib1.so lib2.so ...   (runs bench.py for each experimental build)
for lib in "$@"; do
  UDE_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$lib: value %.3e ms %.1f fwd %.1f bwd %.1f frac_bwd %.3f frac_fwd %.3f e2e %.3e'%(d['value'],d['ms_per_step'],r['fwd_kernel']['ms'],r['bwd_ms'],r['frac'],r['fwd_kernel']['frac'],d['e2e']['value']))
    elif 'Error' in l or 'error' in l: print(l.rstrip())
"
done
