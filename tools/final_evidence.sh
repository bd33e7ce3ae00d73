#!/bin/bash
# End-of-session evidence on one B200 (run under gpurun from the repo root): bench lines of both arms, the launch list of the
# bench command and the film kernels of the Gaussian-film run.  Every ncu pass runs only after the same command exited 0 without ncu.
set -u
O=gpurun_out
python bench.py --steps 10 --warmup 3 > $O/s2_bench_c2.json 2> $O/s2_bench.err || exit 1
python bench.py --impl reference --steps 1 --warmup 0 > $O/s2_bench_c2_reference.json 2>> $O/s2_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --film gaussian > $O/s2_bench_c2_gaussian.json 2>> $O/s2_bench.err
for w in c1 c3 c4 w1 d1; do python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > $O/s2_bench_$w.json 2>> $O/s2_bench.err; done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>> $O/s2_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/s2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/s2_ncu.log 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --film gaussian > /dev/null 2>> $O/s2_bench.err && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:"k_film|k_accumulate" -c 6 --csv --log-file $O/s2_film_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --film gaussian > $O/s2_ncu_film.log 2>&1
tail -2 $O/s2_bench.err
