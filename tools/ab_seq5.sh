# development: GPU parity tests under both two-phase sequence-stage variants, then the stage times
for v in 3 4; do ZN_SEQ=$v python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2 | sed "s/^/ZN_SEQ=$v /"; done
bash tools/ab_seq4.sh
