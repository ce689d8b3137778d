"""wrenc_b200 — B200-native all-intra RD search for the wrenc H.266/VVC encoder (C-ABI library + ctypes host mirror)."""
from .encoder import (DUAL_TREE_CHROMA, DUAL_TREE_LUMA, EXPORTS, LIB_PATH, RECORD_DTYPE, SINGLE_TREE, SearchEncoder,  # noqa: F401
                      WrencB200Error, WrencB200Full, assemble_vvc, header_rbsp, load_library, write_nal, write_parameter_sets, write_picture)
from .synth import random_frame, synth_frame  # noqa: F401
