for w in "" 1; do
for ww in "" ; do
B200_LU_WIDE=$w timeout 600 python -c "
import json, os, bench
r = bench.bench_rsfq(0)
print('WIDE=' + os.environ.get('B200_LU_WIDE',''), json.dumps({k: r[k] for k in ('value','ticks','load_ms_per_tick','solve_ms_per_tick')}), r['parity']['counts_identical'], r['parity']['waves_within_tol'])
" 2>&1 | tail -1
done; done
for W in 4 16 32; do
B200_LU_WIDE=1 B200_LUW_WARPS=$W timeout 600 python -c "
import json, os, bench
r = bench.bench_rsfq(0)
print('WIDE=1 W=' + os.environ.get('B200_LUW_WARPS',''), json.dumps({k: r[k] for k in ('value','ticks','load_ms_per_tick','solve_ms_per_tick')}), r['parity']['counts_identical'], r['parity']['waves_within_tol'])
" 2>&1 | tail -1
done
