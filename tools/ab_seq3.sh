# development: phase split of the two-phase sequence stage + ncu captures of its kernels
B="--workload realtext --steps 2 --warmup 1 --no-cpu --sustain 0 --no-compress"
for v in 3 4; do
  ZN_SEQ=$v ZN_ZPROF=1 ZN_ZPROF_SEQ1=1 python bench.py $B 2>&1 >/dev/null | grep zpipe | tail -1 | sed "s/^/seq=$v: /" | cut -c1-10,100-360
done
ZN_SEQ=3 ncu --set full --clock-control none --import-source on -k regex:"k_zseq1|k_zseq2" -c 2 -f -o gpurun_out/r2_zseq2p_smem python bench.py $B > gpurun_out/ncu_f.log 2>&1
ZN_SEQ=4 ncu --set full --clock-control none --import-source on -k regex:"k_zseq1" -c 1 -f -o gpurun_out/r2_zseq2p_glob python bench.py $B > gpurun_out/ncu_g.log 2>&1
ls -la gpurun_out/r2_zseq2p*
