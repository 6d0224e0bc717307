for L in "" $(ls patchmixturekriging_b200/libpmk_b200_*.so); do PMK_LIB=${L:+$PWD/$L} python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('${L:-product}', 'fit', round(d['ms_fit'],2), 'chol', round(d['phases']['fit_chol_ms'],2), 'pairs', round(d['phases']['query_pairs_ms'],1))"; done
