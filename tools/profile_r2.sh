#!/bin/bash
# Round-2 profiling pass (run under gpurun on ONE GPU): ncu --set full of the dominant kernel of every config, and the
# launch list of the bench command.  Each profiled command first runs plain and must exit 0.
set -u
mkdir -p gpurun_out
prof() {  # name, kernel regex, args of tools/prof_one.py
  local name=$1 re=$2; shift 2
  timeout 120 python tools/prof_one.py "$@" > gpurun_out/plain_$name.log 2>&1 &&
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$re -s 2 -c 1 -f -o gpurun_out/r2_ncu_$name \
      python tools/prof_one.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
}
prof r2c4096 pow2_r2c_stream rfft 4096 65536
prof c2c4096 pow2_c2c_stream cfft 4096 65536
prof mix1001_cosq mix_stream cosq 1001 32768
prof mix1000_cost mix_stream cost 1001 32768
prof mix999_cost mix_stream cost 1000 32768
prof mix1002_sint mix_stream sint 1001 32768
timeout 250 python bench.py --steps 2 --no-cpu-baseline --no-configs > gpurun_out/plain_bench.log 2>&1 &&
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2_ncu_launches_bench.csv python bench.py --steps 2 --no-cpu-baseline --no-configs > gpurun_out/ncu_bench.log 2>&1
echo "bench launch list rc=$?"
