"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the CPU
oracle). CPU: the oracle still reproduces them. GPU: the CUDA product reproduces them bit-exactly
through the C ABI — this needs neither /root/reference nor a fresh oracle run."""
import os

import numpy as np
import pytest

from khronos_b200 import capi
import harness as hs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_camera():
    from khronos_b200 import synthetic as syn
    return syn.make_camera(80, 60, 40.0, 40.0, max_range=2.5)


def check_blocks(g, b: capi.Blocks):
    np.testing.assert_array_equal(g["b_block_index"], b.block_index)
    for k in ("block_flags", "last_observed", "last_occupied", "ever_free", "active", "to_remove",
              "semantic_label", "semantic_empty"):
        np.testing.assert_array_equal(g["b_" + k], getattr(b, k), err_msg=k)
    np.testing.assert_array_equal(g["b_distance"].view(np.uint32), b.distance.view(np.uint32))
    np.testing.assert_array_equal(g["b_weight"].view(np.uint32), b.weight.view(np.uint32))
    np.testing.assert_array_equal(g["b_lik_values"], b.semantic_likelihoods[b.semantic_empty == 0])


def run_fusion_case(lib, prefix):
    g = np.load(os.path.join(GOLD, "fusion.npz"))
    h = hs.make_handle(lib, prefix, cam=golden_camera())
    frames = list(zip(g["depth"], g["label"]))
    frames = [(np.ascontiguousarray(d), np.ascontiguousarray(l)) for d, l in frames]
    stats = hs.run_fusion(h, frames, list(g["poses"]), [int(s) for s in g["stamps"]], tracking=True)
    np.testing.assert_array_equal(g["stats"], np.array([[s[k] for k in sorted(s)] for s in stats], np.int64))
    check_blocks(g, h.export_blocks())


def run_dynamic_case(lib, prefix):
    g = np.load(os.path.join(GOLD, "dynamic.npz"))
    mot = capi.default_motion_config(min_cluster_size=4, min_separation_distance=2.0)
    h = hs.make_handle(lib, prefix, cam=golden_camera(), mot_cfg=mot)
    for i in range(len(g["stamps"])):
        d, l = np.ascontiguousarray(g["depth"][i]), np.ascontiguousarray(g["label"][i])
        T, st = g["poses"][i], int(g["stamps"][i])
        img, ns, nc = h.detect_motion(h.make_frame(d, T, st, label=l))
        assert (ns, nc) == tuple(g["seeds_clusters"][i]), i
        np.testing.assert_array_equal(img.astype(np.uint8), g["dynamic_image"][i])
        h.integrate_frame(h.make_frame(d, T, st, label=l, mask=img))
        h.update_tracking(st)
    check_blocks(g, h.export_blocks())


def run_colour_case(lib, prefix):
    g = np.load(os.path.join(GOLD, "colour.npz"))
    h = hs.make_handle(lib, prefix, cam=golden_camera())
    for i in range(len(g["stamps"])):
        d, l, c = (np.ascontiguousarray(g[k][i]) for k in ("depth", "label", "color"))
        h.integrate_frame(h.make_frame(d, g["poses"][i], int(g["stamps"][i]), label=l, color=c if g["has_color"][i] else None))
    b = h.export_blocks()
    check_blocks(g, b)
    np.testing.assert_array_equal(g["b_color"], b.color)


def run_mesh_case(lib, prefix):
    """Mesh + map checksum of the fusion fixture's map (tests/golden/mesh.npz)."""
    g, m = np.load(os.path.join(GOLD, "fusion.npz")), np.load(os.path.join(GOLD, "mesh.npz"))
    h = hs.make_handle(lib, prefix, cam=golden_camera())
    frames = [(np.ascontiguousarray(d), np.ascontiguousarray(l)) for d, l in zip(g["depth"], g["label"])]
    hs.run_fusion(h, frames, list(g["poses"]), [int(s) for s in g["stamps"]], tracking=True)
    assert tuple(int(x) for x in m["checksum"]) == h.map_checksum()
    bi, off, pts, col, lab = h.generate_mesh(False, False)
    np.testing.assert_array_equal(m["block_index"], bi)
    np.testing.assert_array_equal(m["offsets"], off)
    np.testing.assert_array_equal(m["points_bits"], pts.view(np.uint32))
    np.testing.assert_array_equal(m["labels"], lab.astype(np.uint8))


def test_oracle_matches_golden_mesh(oracle_lib):
    run_mesh_case(oracle_lib, "ko_")


def test_oracle_matches_golden_fusion(oracle_lib):
    run_fusion_case(oracle_lib, "ko_")


def test_oracle_matches_golden_dynamic(oracle_lib):
    run_dynamic_case(oracle_lib, "ko_")


def test_oracle_matches_golden_colour(oracle_lib):
    run_colour_case(oracle_lib, "ko_")


def test_oracle_is_thread_count_invariant(oracle_lib):
    """Determinism under block-order permutation: 1 thread vs all threads."""
    g = np.load(os.path.join(GOLD, "fusion.npz"))
    h = hs.make_handle(oracle_lib, "ko_", cam=golden_camera(),
                       integ_cfg=capi.default_integrator_config(num_threads=1),
                       trk_cfg=capi.default_tracking_config(num_threads=1))
    frames = [(np.ascontiguousarray(d), np.ascontiguousarray(l)) for d, l in zip(g["depth"], g["label"])]
    hs.run_fusion(h, frames, list(g["poses"]), [int(s) for s in g["stamps"]], tracking=True)
    check_blocks(g, h.export_blocks())


@pytest.mark.gpu
def test_product_matches_golden_mesh(product_lib):
    run_mesh_case(product_lib, "kb_")


@pytest.mark.gpu
def test_product_matches_golden_fusion(product_lib):
    run_fusion_case(product_lib, "kb_")


@pytest.mark.gpu
def test_product_matches_golden_dynamic(product_lib):
    run_dynamic_case(product_lib, "kb_")
