"""oracle/ -- TEST INFRASTRUCTURE, not product code.

CPU restatement of the reference hot path (feiyuhuahuo/Yolact_minimal @ d920c05):
  * postprocess_np.py : numpy restatement of utils/output_utils.py (nms, fast_nms,
                        traditional_nms, after_nms), utils/box_utils.py (box_iou, crop,
                        sanitize_coordinates, make_anchors) and cython_nms.pyx.
  * forward_torch.py  : plain-torch fp32 functional restatement of modules/yolact.py +
                        modules/resnet.py (floating-point kernel reference).
  * hard_nms.c        : plain-C restatement of cython_nms.pyx:24-74.
  * train_np.py       : numpy restatement of the training targets and losses (match / encode, OHEM mining, the four
                        losses, mask_iou), stage by stage -- the checker for native training kernels.
  * synth.py          : deterministic, library-independent synthetic input generator shared
                        by the golden generator and the tests.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import anything from this package, and only as the checker.  The product package
(yolact_minimal_b200/) never imports it and fails loudly when its CUDA library is missing.

Parity status: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the build container
by tests/golden/make_golden.py (imports /root/reference) and committed under tests/golden/.
"""
