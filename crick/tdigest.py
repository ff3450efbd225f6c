from crick_b200.tdigest import CENTROID_DTYPE, TDigest, TDigestGroup, grouped_update  # noqa: F401
