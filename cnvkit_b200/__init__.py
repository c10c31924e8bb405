"""cnvkit_b200 -- B200-native (sm_100a) implementation of CNVkit's numeric hot path.

Host-side mirror of the reference's Python seam:

    cnvkit_b200.smoothing.rolling_median      <- cnvlib/smoothing.py:77
    cnvkit_b200.fix.do_fix                    <- cnvlib/fix.py:20
    cnvkit_b200.reference.do_reference        <- cnvlib/reference.py:88
    cnvkit_b200.segmentation.do_segmentation  <- cnvlib/segmentation/__init__.py:31

All numerics run in hand-written CUDA kernels behind the C ABI of
``libcnvkit_b200.so`` (include/cnvkit_b200.h).  There is no CPU fallback.
"""
__version__ = "0.1.0"
