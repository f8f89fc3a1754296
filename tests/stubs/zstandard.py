"""Test-only stand-in for the `zstandard` package (absent from this image), for running the UNMODIFIED reference client
`baseline/_ref/pmpc/remote.py` (it does `import zstandard` at module level) against this repository's worker: the three
names the reference uses, on top of the system's libzstd (pmpc_b200.remote._LibZstd).  Only on PYTHONPATH of the
subprocess tests/test_remote_wire.py starts."""
from pmpc_b200.remote import _LibZstd

_codec = _LibZstd()
compress, decompress, ZstdError = _codec.compress, _codec.decompress, _codec.ZstdError
