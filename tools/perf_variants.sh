#!/bin/bash
# ad-hoc: time the turbo kernel under several tuning knobs (one JSON line each), K = 5824, 53 248 blocks
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
export SRSUE_TURBO_PHASE_DELAY=40000
for mode in "4 2 30.0" "4 2 2.0" "4 0 2.0"; do
  tag=$(echo $mode | tr ' ' '_')
  run "generic_$tag" SRSUE_TURBO_GENERIC=1 -- 5824 53248 $mode
  run "t26_$tag" A=1 -- 5824 53248 $mode
  run "t26_d60k_$tag" SRSUE_TURBO_PHASE_DELAY=60000 -- 5824 53248 $mode
  run "t26_d30k_$tag" SRSUE_TURBO_PHASE_DELAY=30000 -- 5824 53248 $mode
done
run "K6144" A=1 -- 6144 53248 4 0 2.0
run "K40" A=1 -- 40 1000000 4 0 2.0
python -c "
import json
for l in open('$out'):
    d=json.loads(l); r=d['r']
    print(d['label'], None if r is None else (round(r['ms'],3), round(r['avg_iters'],3), round(r['int16_tops'],2), r['launch']))
"
tail -2 gpurun_out/perf_variants.err
