"""Writes tests/golden/spark_hash_vectors.json.  The expected values are the reference's stored vector
(BucketUnionTest.scala:101-123) and Spark's published hash values; the script recomputes them with the oracle and
refuses to write if the oracle disagrees."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

VECTORS = {
    "bucket_union_test": {"keys_int32": [2, 3], "num_partitions": 10, "partitions": [4, 1],
                          "per_partition_key_sums_of_union": [0, 6, 0, 0, 4, 0, 0, 0, 0, 0]},
    "hash_long_seed42": {"0": -1670924195, "1": -1712319331, "2": -797927272, "3": 519220707, "-1": -939490007},
    "pmod200_of_hash_long": {"0": 5, "1": 69, "2": 128, "3": 107, "-1": 193},
    # Spark 3.1.1 Murmur3Hash ExpressionDescription (the functions documentation): bytes + int + seed fold
    "spark_docs_hash_example": {"sql": "SELECT hash('Spark', array(123), 2)", "seed": 42, "result": -1321691492},
}

if __name__ == "__main__":
    assert O.bucket_ids([np.array([2, 3], dtype=np.int32)], 10).tolist() == VECTORS["bucket_union_test"]["partitions"]
    for k, h in VECTORS["hash_long_seed42"].items():
        assert O.lib().hso_hash_long(int(k), 42) == h, k
        assert int(O.bucket_ids([np.array([int(k)], dtype=np.int64)], 200)[0]) == VECTORS["pmod200_of_hash_long"][k]
    L = O.lib()
    assert L.hso_hash_int(2, L.hso_hash_int(123, L.hso_hash_bytes(b"Spark", 5, 42))) == VECTORS["spark_docs_hash_example"]["result"]
    with open(os.path.join(os.path.dirname(__file__), "spark_hash_vectors.json"), "w") as f:
        json.dump(VECTORS, f, indent=2)
    print("ok")
