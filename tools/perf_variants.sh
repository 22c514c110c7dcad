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
for mode in "4 2 2.0" "4 2 30.0"; do
  tag=$(echo $mode | tr ' ' '_')
  run "g1_$tag" SRSUE_TURBO_GROUPS=1 -- 5824 53248 $mode
  run "g2_d0_$tag" SRSUE_TURBO_GROUPS=2 SRSUE_TURBO_PHASE_DELAY=0 -- 5824 53248 $mode
  run "g2_d10k_$tag" SRSUE_TURBO_GROUPS=2 SRSUE_TURBO_PHASE_DELAY=10000 -- 5824 53248 $mode
  run "g2_d20k_$tag" SRSUE_TURBO_GROUPS=2 SRSUE_TURBO_PHASE_DELAY=20000 -- 5824 53248 $mode
  run "g2_d40k_$tag" SRSUE_TURBO_GROUPS=2 SRSUE_TURBO_PHASE_DELAY=40000 -- 5824 53248 $mode
done
cat $out
