#!/bin/bash
# ad-hoc: time the turbo kernel under several tuning knobs / build variants (one JSON line each), K = 5824, 53 248 blocks
out=${1:-gpurun_out/perf_variants.jsonl}
: > $out
run() {  # label, env..., -- args
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo -n "{\"label\": \"$label\", \"r\": " >> $out
  env "${envs[@]}" python tools/perf_turbo.py "$@" >> $out 2>>gpurun_out/perf_variants.err || echo "null" >> $out
  sed -i '$ s/$/}/' $out
}
for mode in "4 2 30.0" "4 2 2.0" "4 0 2.0"; do
  tag=$(echo $mode | tr ' ' '_')
  run "base_$tag" A=1 -- 5824 53248 $mode
  for v in $VARIANTS; do
    run "${v}_$tag" SRSUE_GPU_LIB=build/variants/libsrsue_gpu_$v.so -- 5824 53248 $mode
  done
done
python -c "
import json
for l in open('$out'):
    d=json.loads(l); r=d['r']
    print(d['label'], None if r is None else (round(r['ms'],3), round(r['avg_iters'],3), round(r['int16_tops'],2), r['launch']))
"
tail -2 gpurun_out/perf_variants.err
