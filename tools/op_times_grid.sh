#!/bin/bash
# Operation-time sweep of the reference (scripts/run_fhn_model_noiseless_obs_experiments.sh:16-22) on the batched
# kernels: one line per grid point into gpurun_out/$1 (default r2_op_times_grid.jsonl).  Grid points the per-thread
# block algebra cannot hold (num_obs_per_subseq >= 20: more than 16 constraint rows per block) are skipped.
out=gpurun_out/${1:-r2_op_times_grid.jsonl}
: > $out
for T in 25 50 100 200 400; do CPU=0 T=$T S=25 R=5 NSTATES=1024 timeout 300 python tools/op_times.py >> $out 2>> $out.err; done
for S in 50 100 200 400; do CPU=0 T=100 S=$S R=5 NSTATES=1024 timeout 300 python tools/op_times.py >> $out 2>> $out.err; done
for R in 2 10; do CPU=0 T=100 S=25 R=$R NSTATES=1024 timeout 300 python tools/op_times.py >> $out 2>> $out.err; done
CPU=1 T=100 S=25 R=5 NSTATES=1024 timeout 600 taskset -c 0 python tools/op_times.py > gpurun_out/r2_op_times.json 2>> $out.err
wc -l $out; tail -c 400 gpurun_out/r2_op_times.json
