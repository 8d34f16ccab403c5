# development (run under gpurun): the sequence stage of the zstd pipeline, form by form.
#   ZN_SEQ = 0 one-pass, tables in global memory (k_zseq_g)        1 one-pass, tables in shared memory (k_zseq)
#            3 two-phase, phase-1 tables in shared memory (k_zseq1 + k_zseq2; the default for large blobs)
#            4 two-phase, phase-1 tables in global memory (k_zseq1_g + k_zseq2)
# Parity of both two-phase forms first, then the stage times (ZN_ZPROF; ZN_ZPROF_SEQ1 adds the end of phase 1).
for v in 3 4; do ZN_SEQ=$v python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1 | sed "s/^/ZN_SEQ=$v /"; done
B="--steps 2 --warmup 1 --no-cpu --sustain 0 --no-compress"
for w in realtext realsmall; do for v in 0 3 4; do
  ZN_SEQ=$v ZN_ZPROF=1 ZN_ZPROF_SEQ1=1 python bench.py --workload $w $B 2>&1 >/dev/null | grep zpipe | tail -1 | sed "s/^/seq=$v $w: /" | cut -c1-20,100-360
done; done
