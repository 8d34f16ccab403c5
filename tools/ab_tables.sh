# development: warp-parallel FSE table construction (k_ztables) — parity first, then the stage times of the small-file corpus
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
ZN_ZPROF=1 python bench.py --workload realsmall --steps 2 --warmup 1 --no-cpu --sustain 0 --no-compress 2>&1 >/dev/null | grep zpipe | tail -1 | cut -c100-330
