#!/bin/bash
# DRAM traffic / L2 hit rate of the headline kernel under ncu for a list of (lib variant, flags, keep, ctas_per_sm) settings.
# usage: tools/sv_traffic.sh out.txt "lib:flags:keep:cps" ...
out=$1; shift
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum
: > $out
for c in "$@"; do
  IFS=: read lib flags keep cps <<< "$c"
  if [ -n "$lib" ]; then export HPMPC_B200_LIB=$PWD/hpmpc_b200/lib/variants/libhpmpc_b200_$lib.so; else unset HPMPC_B200_LIB; fi
  export HPMPC_B200_SV_FLAGS=$flags HPMPC_B200_SV_KEEP=$keep
  echo "== lib=${lib:-default} flags=$flags keep=$keep cps=$cps" >> $out
  ncu --metrics $M --clock-control none -k regex:hbk_ric_sv -s 2 -c 1 python tools/prof_sv.py 65536 $cps 0 1 2>&1 | grep -E "dram__bytes|hit_rate|gpu__time|lts__t_sectors|solves/s" >> $out
done
cat $out
