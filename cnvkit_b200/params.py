"""Constants of the hot path -- same values as cnvlib/params.py:3-19."""
MIN_REF_COVERAGE = -5.0
MAX_REF_SPREAD = 1.0
NULL_LOG2_COVERAGE = -20.0
GC_MIN_FRACTION = 0.3
GC_MAX_FRACTION = 0.7
INSERT_SIZE = 250
IGNORE_GENE_NAMES = ("-", ".", "CGH")
ANTITARGET_NAME = "Antitarget"
ANTITARGET_ALIASES = (ANTITARGET_NAME, "Background")
