"""hyperspace_b200 -- B200-native covering-index engine behind the Hyperspace API surface.

The data path (Parquet decode, Spark-compatible bucket hashing, partition, sort, Parquet encode, filter scan and
bucket-aligned merge join) runs in hand-written sm_100a CUDA inside ``lib/libhs_gpu.so`` (C ABI: ``include/hs_gpu.h``).
This package holds the ctypes binding (``_native``) and the host-side mirror of the reference's Python API
(``python/hyperspace/hyperspace.py`` in microsoft/hyperspace).
"""
__version__ = "0.1.0"
