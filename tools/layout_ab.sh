#!/bin/bash
# A/B of the packed node order (RDF_PACK_LAYOUT=heap|blocks) x L2 fetch granularity on the divergent and the coherent workloads
for lay in heap blocks; do for gran in 0 64 128; do
  export RDF_PACK_LAYOUT=$lay; if [ $gran = 0 ]; then unset RDF_L2_FETCH; else export RDF_L2_FETCH=$gran; fi
  for w in cfg5-noise cfg5 cfg3-noise cfg3; do
    fr=""; st=10; if [ $w = cfg3 ] || [ $w = cfg3-noise ]; then fr="--frames 512"; st=3; fi
    v=$(python bench.py --no-extras --steps $st $fr --workload $w 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['parity_checked'])")
    echo "$lay fetch=$gran $w $v"
  done
done; done
