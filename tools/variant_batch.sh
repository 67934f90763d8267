#!/bin/bash
# Runs tools/kbench.py (16,384 chains) for each library variant given on the command line and, with NCU=1, an ncu
# metrics pass (DRAM bytes, duration, local-memory sectors) over two k_leapfrog launches of the same command.
# usage: tools/variant_batch.sh TAG lib1.so lib2.so ...   (results in gpurun_out/TAG_*.{json,csv})
tag=$1; shift
for lib in "$@"; do
  name=$(basename $lib .so)
  MMD_B200_LIB=$lib NCH=${NCH:-16384} TAG=$name python tools/kbench.py > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  cat gpurun_out/${tag}_${name}.json
  if [ "${NCU:-0}" = "1" ]; then
    MMD_B200_LIB=$lib NCH=${NCH:-16384} BURN=6 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum \
      --clock-control none -k regex:k_leapfrog -s 7 -c 2 --csv --log-file gpurun_out/${tag}_${name}_ncu.csv python tools/kbench.py > /dev/null 2>&1
    grep -v "^==" gpurun_out/${tag}_${name}_ncu.csv | awk -F'","' 'NR>1 {gsub(/"/,"",$NF); print $(NF-2), $NF}' | tr '\n' ';'; echo
  fi
done
