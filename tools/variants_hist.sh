#!/bin/bash
# A/B run of librdf_b200.so variants (tools/build_variant.sh) on the cfg4 level phases: usage tools/variants_hist.sh "0,8,12" base u8 c2 ...
LEVELS="$1"; shift
for v in "$@"; do
  if [ "$v" = base ]; then unset RDF_B200_LIB; else export RDF_B200_LIB="$PWD/build/variants/$v/librdf_b200.so"; fi
  echo "== $v"
  python tools/bench_train_phases.py --levels "$LEVELS" $EXTRA 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print(d['level'], 'hist', d['ms']['hist'], 'pick', d['ms']['pick_best'], d['records_md5'][:8])
    except Exception:
        print(l[:200])
"
done
