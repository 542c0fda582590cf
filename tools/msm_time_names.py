"""Short names of the hk_ctx_set_option knobs shared by the tuning helpers (msm_time.py, msm_once.py, bench.py --opt)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from hekaton_system_b200 import backend as B

NAMES = {"waves": B.OPT_ACCUM_WAVES, "smin": B.OPT_ACCUM_SEG_MIN, "fold": B.OPT_FOLD_SEG, "chunk": B.OPT_REDUCE_CHUNK,
         "gsort": B.OPT_SORT_GLOBAL_ATOMICS, "g2pair": B.OPT_G2_LANE_PAIR, "quad": B.OPT_QUAD_MAX_THREADS,
         "noresolve": B.OPT_NO_FOLD_RESOLVE}
