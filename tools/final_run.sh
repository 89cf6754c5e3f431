#!/bin/bash
# Round-end evidence run on one GPU box: full GPU test suite, smoke, the bench lines (this repo's
# arm, the reference's CPU arm, the library GPU arm) and the ncu passes tools/make_profiles.py reads.
# usage: tools/final_run.sh <tag>        (writes gpurun_out/*_<tag>.*)
tag=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/tests_$tag.log; tail -1 gpurun_out/tests_$tag.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_n1_$tag.json 2> gpurun_out/bench_n1_$tag.err; tail -c 600 gpurun_out/bench_n1_$tag.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$tag.json 2>/dev/null
python bench.py --impl library --steps 3 --warmup 2 > gpurun_out/bench_library_$tag.json 2>/dev/null
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/kernels_$tag.csv python bench.py --steps 1 --warmup 3 \
    --graph off --no-cpu-baseline --no-profile --no-secondary --no-library-baseline > gpurun_out/kernels_bench_$tag.log 2>&1
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/fwd256_$tag.csv python tools/fwd256.py 256 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 40 -c 20 -f \
    -o gpurun_out/prof_fwd256_$tag python tools/fwd256.py 256 > /dev/null 2>&1
ls -la gpurun_out/*_$tag.* | head -20
