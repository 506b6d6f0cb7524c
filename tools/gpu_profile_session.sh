#!/bin/bash
# One profiling session on the GPU box (run through gpurun): every capture is reduced to text ON THE BOX (gpurun copies back at most
# 64 MiB, .ncu-rep files are far larger) -- summaries via tools/ncu_summary.py, stall tables via tools/ncu_stalls.py.
#   gpurun --timeout 1500 -- 'bash tools/gpu_profile_session.sh'
O=gpurun_out
mkdir -p $O
reduce() {   # reduce <name>: $O/<name>.ncu-rep -> $O/<name>_summary.txt, $O/<name>_stalls.txt ; the report itself is dropped
    python tools/ncu_summary.py $O/$1.ncu-rep > $O/$1_summary.txt 2>&1
    ncu -i $O/$1.ncu-rep --page source --csv > /tmp/$1_src.csv 2>/dev/null && python tools/ncu_stalls.py /tmp/$1_src.csv 10 > $O/$1_stalls.txt 2>&1
}
# (SKIP="lin user" leaves out the sections whose kernels have not changed since their last capture)
# 1. linear Gram kernel: ring depth / steps per group
if [[ " $SKIP " != *" lin "* ]]; then
: > $O/r2_lin_sweep.txt
for un in 2 4; do for st in 3 4 6; do
  echo -n "UN=$un ST=$st " >> $O/r2_lin_sweep.txt
  GSA_LIN_UN=$un GSA_LIN_ST=$st python tools/bench_configs.py --only c4 --no-y 2>&1 | sed 's/.*"kernel_ms": \([0-9.]*\).*"frac_hbm": \([0-9.]*\).*/kernel_ms \1 frac \2/' >> $O/r2_lin_sweep.txt
done; done
cat $O/r2_lin_sweep.txt
fi
# 2. the headline kernel: full capture at N = 2^24 (27 GB of design), traffic json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_product_kernel -c 1 -f -o $O/r2_prod \
    python bench.py --no-cpu --no-e2e --no-configs --no-gen --steps 1 --warmup 1 --log2n 24 > $O/r2_ncu_prod.log 2>&1
python tools/make_traffic_json.py $O/r2_prod.ncu-rep 24 > $O/r2_traffic.json 2>> $O/r2_ncu_prod.log
reduce r2_prod; rm -f $O/r2_prod.ncu-rep
cat $O/r2_traffic.json
# 3. user models: generic kernel (FP64 pipe evidence) and separable kernels
if [[ " $SKIP " != *" user "* ]]; then
python tools/user_model_probe.py > $O/r2_user_models.jsonl 2> $O/r2_user_models.err
for route in generic generic_lto separable separable_lto; do for dd in 20 100; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gsa_user --launch-skip 1 -c 1 -f -o $O/r2_u_${route}_$dd \
      python tools/user_model_probe.py --only $route,$dd > $O/r2_ncu_user.log 2>&1
  reduce r2_u_${route}_$dd; rm -f $O/r2_u_${route}_$dd.ncu-rep
done; done
cat $O/r2_u_*_summary.txt | grep -E "^##|time=" | cut -c1-260
fi
# 4. every other hot kernel
timeout 600 ncu --set full --clock-control none --import-source on -f -o $O/r2_prof python tools/profile_run.py > $O/r2_ncu_prof.log 2>&1
reduce r2_prof; rm -f $O/r2_prof.ncu-rep
# 5. launch list of the default bench command (after it has run once without ncu)
python bench.py --steps 2 --warmup 1 --no-cpu > $O/r2_b_plain.json 2> $O/r2_b_plain.err && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/r2_ncu_launch.log 2>&1
tail -2 $O/r2_ncu_launch.log
du -sh $O
