for v in base a64 a64b minb3 t128 a128; do echo "== $v"; NCT_LIB=neural-color-transfer_b200/variants/libnct_$v.so python tools/pm_levels.py 700 2>&1 | tail -6 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    print(d.get('level','tot'), d.get('C',''), d.get('ms', d.get('total_ms')))"; done
