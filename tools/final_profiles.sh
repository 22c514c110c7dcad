set -x
python bench.py > gpurun_out/bench_final_r02.json 2> gpurun_out/bench_final_r02.err
python bench.py --impl reference > gpurun_out/bench_ref_final_r02.json 2>> gpurun_out/bench_final_r02.err
python bench.py --workload mixed > gpurun_out/bench_mixed_final_r02.json 2>> gpurun_out/bench_final_r02.err
python bench.py --workload mixed --blind > gpurun_out/bench_mixed_blind_final_r02.json 2>> gpurun_out/bench_final_r02.err
python bench.py --workload harq > gpurun_out/bench_harq_final_r02.json 2>> gpurun_out/bench_final_r02.err
python bench.py --workload tm2 > gpurun_out/bench_tm2_final_r02.json 2>> gpurun_out/bench_final_r02.err
python bench.py --batch 1024 --steps 2 --warmup 3 --no-cpu-baseline --no-legs > gpurun_out/plain_chain_final.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final_r02.csv python bench.py --batch 1024 --steps 2 --warmup 3 --no-cpu-baseline --no-legs > gpurun_out/ncu_l_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ofdm_rx|chest_kernel|pdsch_llr|turbo_decode|tdec_deint|tb_assemble" --launch-skip 18 --launch-count 12 -o gpurun_out/prof_chain_final_r02 -f python bench.py --batch 1024 --steps 2 --warmup 3 --no-cpu-baseline --no-legs > gpurun_out/ncu_c_final.log 2>&1
tail -c 600 gpurun_out/bench_final_r02.json; tail -3 gpurun_out/bench_final_r02.err
