"""Digest of a clustering result (every field the parity tests compare), used to pin the BASELINE configurations at
their FULL sizes: the digests of the CPU oracle's results are committed in tests/golden/config_hashes.json (written by
tests/golden/make_config_hashes.py in the build container, where the oracle runs), so that bench.py and the -m gpu
tests can check the CUDA path at full size without executing anything under oracle/ on the GPU box."""
import hashlib
import json
import os

import numpy as np

REC_FIELDS = ("sum", "residue_sum", "rect", "num_holes", "depth", "merged_into", "indices_len", "indices_off",
              "holes_off", "holes_len")
HASH_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config_hashes.json")


def _upd(h, a, dtype):
    h.update(np.ascontiguousarray(a, dtype=dtype).tobytes())


def clusters_digest(res: dict) -> str:
    """labels, slot count, every record field, clusters_output (+ to_color_image when present)."""
    h = hashlib.blake2b(digest_size=16)
    h.update(np.array([res["width"], res["height"], res["num_slots"], len(res["clusters_output"])], dtype=np.uint64).tobytes())
    _upd(h, res["cluster_indices"], np.uint32)
    rec = res["records"]
    for f in REC_FIELDS:
        _upd(h, rec[f], np.int64)
    _upd(h, res["clusters_output"], np.uint32)
    if "color_image" in res:
        _upd(h, res["color_image"], np.uint8)
    return h.hexdigest()


def bin_clusters_digest(res: dict) -> str:
    h = hashlib.blake2b(digest_size=16)
    _upd(h, res["rect"], np.int64)
    for f in ("rect", "points_off", "points_len"):
        _upd(h, res["records"][f], np.int64)
    _upd(h, res["points"], np.int32)
    return h.hexdigest()


def components_digest(comps) -> str:
    """vcb_clusters_components / the oracle's to_image_with_hole(..).to_clusters(false) per output cluster: for every output
    cluster its components in order, each (rect, points)."""
    h = hashlib.blake2b(digest_size=16)
    h.update(np.array([len(comps)], dtype=np.uint64).tobytes())
    for cl in comps:
        h.update(np.array([len(cl)], dtype=np.uint64).tobytes())
        for rect, pts in cl:
            _upd(h, np.array(rect), np.int64)
            _upd(h, pts, np.int32)
    return h.hexdigest()


def perimeters_digest(perims, outputs) -> str:
    """Cluster::perimeter of the output clusters, in clusters_output order"""
    h = hashlib.blake2b(digest_size=16)
    _upd(h, np.asarray(perims)[np.asarray(outputs, dtype=np.int64)], np.int64)
    return h.hexdigest()


def load_hashes() -> dict:
    with open(HASH_FILE) as f:
        return json.load(f)
