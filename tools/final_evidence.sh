#!/bin/bash
# Round-2 evidence on one B200 (run under gpurun from the repo root).  Every ncu pass runs only after the same command has
# exited 0 without ncu; numbers printed under ncu are never bench values.
set -u
O=gpurun_out
mkdir -p $O
Q="--no-cpu-baseline --no-bridge --no-strong-record"
# 1. launch lists (gpu__time_duration.sum per launch) of the bench command, C2 and C4
for w in c2 c4; do
  python bench.py --workload $w --steps 2 --warmup 3 $Q > /dev/null 2>> $O/ev.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_$w.csv \
      python bench.py --workload $w --steps 2 --warmup 3 $Q > $O/ev_ncu_$w.log 2>&1
done
# 2. per-launch DRAM bytes + lane / issue statistics of the traversal and VolPath launches (one step, no bench harness)
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
python tools/profile_step.py --workload c2 --steps 2 > $O/ev_step_c2.json 2>> $O/ev.err && \
ncu --metrics $M --clock-control none -k regex:"k_trace|k_shade|k_anyhit" -c 60 --csv --log-file $O/r02_dram_trace_c2.csv python tools/profile_step.py --workload c2 --steps 1 > $O/ev_ncu2_c2.log 2>&1
python tools/profile_step.py --workload c4 --steps 2 > $O/ev_step_c4.json 2>> $O/ev.err && \
ncu --metrics $M --clock-control none -k regex:"k_vp_" -c 400 --csv --log-file $O/r02_dram_trace_c4.csv python tools/profile_step.py --workload c4 --steps 1 > $O/ev_ncu2_c4.log 2>&1
python tools/profile_step.py --workload u1p --steps 2 > $O/ev_step_u1p.json 2>> $O/ev.err && \
ncu --metrics $M --clock-control none -k regex:"k_trace|k_shade|k_anyhit" -c 160 --csv --log-file $O/r02_dram_trace_u1p.csv python tools/profile_step.py --workload u1p --steps 1 > $O/ev_ncu2_u1p.log 2>&1
# 3. full captures of the dominant kernels, exported as text on the box
ncu --set full --clock-control none --import-source on -k regex:"k_trace<0" -c 1 -o $O/r02_full_trace0 -f python tools/profile_step.py --workload c2 --steps 1 > $O/ev_ncu3.log 2>&1 && \
ncu -i $O/r02_full_trace0.ncu-rep --page details > $O/r02_full_trace0.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_anyhit8" -c 1 -o $O/r02_full_anyhit8 -f python tools/profile_step.py --workload c2 --steps 1 > $O/ev_ncu3b.log 2>&1 && \
ncu -i $O/r02_full_anyhit8.ncu-rep --page details > $O/r02_full_anyhit8.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_vp_track" -s 2 -c 1 -o $O/r02_full_vp_track -f python tools/profile_step.py --workload c4 --steps 1 --spp 16 > $O/ev_ncu4.log 2>&1 && \
ncu -i $O/r02_full_vp_track.ncu-rep --page details > $O/r02_full_vp_track.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_vp_logic" -s 8 -c 1 -o $O/r02_full_vp_vertex -f python tools/profile_step.py --workload c4 --steps 1 --spp 16 > $O/ev_ncu5.log 2>&1 && \
ncu -i $O/r02_full_vp_vertex.ncu-rep --page details > $O/r02_full_vp_vertex.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_shade" -s 1 -c 1 -o $O/r02_full_shade_u1p -f python tools/profile_step.py --workload u1p --steps 1 --spp 8 > $O/ev_ncu6.log 2>&1 && \
ncu -i $O/r02_full_shade_u1p.ncu-rep --page details > $O/r02_full_shade_u1p.txt 2>&1 && \
ncu -i $O/r02_full_shade_u1p.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02_shade_u1p_source.csv.gz
rm -f $O/*.ncu-rep
tail -3 $O/ev.err
