"""zk-odst B200 prover backend — Python harness over the C-ABI library `libzkodst.so`.

The product is the CUDA/C++ library (csrc/, include/zkodst.h); this package only loads it with
ctypes for tests and benchmarks.  There is no CPU fallback: every compute entry point raises
when the library or a CUDA device is missing.
"""
from .binding import (  # noqa: F401
    Context,
    ZkError,
    load_library,
    library_path,
    rows_per_compression,
    min_k,
    layout_hash,
    pinned_debug,
    layout_tables,
    dist_unique_id,
    eip152_validate,
    blake2f_compress,
    blake2b_records,
    dist_range,
    dist_column_block,
    dist_quotient_rows,
)
from .inputs import eip152_record, synthetic_inputs, XorShiftRng, REFERENCE_SEED  # noqa: F401
