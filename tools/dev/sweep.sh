#!/bin/bash
# development aid: run the two kNN probes against every variant library given on the command line
for v in "$@"; do
  lib=rolo_b200/lib/variants/$v.so; [ "$v" = base ] && lib=rolo_b200/lib/librolo_gicp_b200.so
  sub=$(ROLO_KNN_STATS= timeout 100 python tools/dev/with_lib.py $lib tools/dev/knn_probe_submap.py 77 2>&1 | grep "^target" | tail -1)
  ben=$(timeout 200 python tools/dev/with_lib.py $lib bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['stages_ms_per_step']['knn_kernels_ms'],2))")
  echo "$v | submap: $sub | batch: $ben"
done
