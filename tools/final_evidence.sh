#!/bin/bash
# Round-end evidence pass on the GPU box: full GPU test suite, smoke, bench (both arms), timelines, sweep, side-kernel
# timings, ncu launch list and full captures.  Everything lands in gpurun_out/ (copy what should be judged into profiles/).
tag=${1:-final}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$tag.log 2>&1; python tools/show_bench.py gpurun_out/bench_$tag.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.log 2>&1; tail -1 gpurun_out/bench_ref_$tag.log | cut -c1-300
python bench.py --steps 20 --warmup 5 --experts literal --no-cpu-baseline --sampling-batches 0 > gpurun_out/bench_literal_$tag.log 2>&1; python tools/show_bench.py gpurun_out/bench_literal_$tag.log
MOE_LIB_VARIANT=trace timeout 100 python tools/trace_fused.py > gpurun_out/trace_$tag.log 2>&1
python tools/sweep_fused.py > gpurun_out/sweep_$tag.log 2>&1; cat gpurun_out/sweep_$tag.log
python tools/time_aux.py > gpurun_out/time_aux_$tag.log 2>&1; cat gpurun_out/time_aux_$tag.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sampling-batches 0 --no-aux > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sampling-batches 0 --no-aux > gpurun_out/ncu_bench_$tag.log 2>&1
for shape in "320 8192" "1280 512"; do
  set -- $shape
  python tools/ncu_one_layer.py $1 $2 > /dev/null 2>&1 && ncu --set full --import-source on --clock-control none -k regex:ffn_fused --launch-skip 4 --launch-count 2 -o gpurun_out/prof_${tag}_d$1 -f python tools/ncu_one_layer.py $1 $2 > /dev/null 2>&1
done
python tools/run_aux_kernels.py > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:"router_multi|hist_accumulate|perm_" -o gpurun_out/prof_${tag}_aux -f python tools/run_aux_kernels.py > /dev/null 2>&1
ls gpurun_out/*${tag}*
