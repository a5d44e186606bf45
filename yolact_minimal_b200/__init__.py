"""yolact_minimal_b200 -- B200-native (sm_100a) YOLACT hot path behind the reference's Python API.

Import points kept from feiyuhuahuo/Yolact_minimal:
    from yolact_minimal_b200.modules.yolact import Yolact
    from yolact_minimal_b200.utils.output_utils import nms, after_nms
    from yolact_minimal_b200.cython_nms import nms as cnms
    from yolact_minimal_b200.config import get_config
All compute goes through libyolact_b200.so (include/yolact_b200.h); there is no CPU fallback.
"""
__version__ = '0.1.0'
