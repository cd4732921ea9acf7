"""B200-native cross-view pose-refinement engine (drop-in for the hot path of YujiaoShi/HighlyAccurate).

    from highlyaccurate_b200.models_kitti import LM_S2GP, loss_func
    from highlyaccurate_b200.models_ford import LM_S2GP_Ford

mirror `models_kitti.py` / `models_ford.py` of the reference; all arithmetic of the path runs in
libha_b200.so (hand-written sm_100a CUDA behind the C ABI in include/ha_b200.h).
"""
__version__ = "0.1.0"
