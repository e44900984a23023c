"""Shared checks of the reference-API mirror (visioncortex_b200/api.py) against the independent Python restatement of the
reference (oracle/vc_restate.py): used on the host simulator (CPU) and on the GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import vc_restate as R  # noqa: E402
from visioncortex_b200 import api  # noqa: E402


def check_runner_api(ctx, img: np.ndarray, **kw):
    """Runner::start() / tick() / progress() / result(), Clusters / ClustersView / Cluster queries."""
    h, w = img.shape[:2]
    cfg = api.RunnerConfig(**{k: (api.Color(*v) if k == "key_color" else v) for k, v in kw.items()})
    ref = R.Builder(R.Config(**kw), [tuple(int(v) for v in p) for p in img.reshape(-1, 4)], w, h).run()

    runner = api.Runner.new(cfg, api.ColorImage.from_array(img))
    runner._ctx = ctx
    ib = runner.start()
    assert ib.progress() == 0
    assert ib.tick() is False          # the whole computation
    assert ib.progress() == 100
    assert ib.tick() is True
    view = ib.view()
    clusters = ib.result()
    assert ib.progress() == 0          # builder.rs:131-135: the builder was taken
    assert clusters.output_len() == len(ref.clusters_output)
    assert [int(v) for v in view.clusters_output] == ref.clusters_output
    assert [int(v) for v in view.cluster_indices] == ref.cluster_indices
    assert len(view.clusters) == len(ref.clusters)
    for k, (got, want) in enumerate(zip(view.clusters, ref.clusters)):
        assert [int(v) for v in got.indices] == want.indices, k
        assert [int(v) for v in got.holes] == want.holes, k
        assert (got.num_holes, got.depth, got.merged_into) == (want.num_holes, want.depth, want.merged_into)
        assert (got.rect.left, got.rect.top, got.rect.right, got.rect.bottom) == want.rect.tup()
        assert got.area() == want.area()
    # queries on the output clusters
    for idx, got in zip(view.clusters_output, view.iter()):
        want = ref.clusters[int(idx)]
        assert got is view.get_cluster(int(idx))
        assert got.color().tuple() == want.sum.average() and got.residue_color().tuple() == want.residue_sum.average()
        assert got.neighbours(view) == ref.neighbours_internal(want)
        bits, bw, bh = R.to_image_with_hole(want, w, True)
        im = got.to_image(view)
        assert (im.width, im.height) == (bw, bh) and bytes(im.pixels.astype(np.uint8)) == bytes(bits)
        assert got.perimeter(view) == R.image_boundary_count(bits, bw, bh)
    # the same queries computed on the device from labels + merge forest + records (no pools): every cluster that holds pixels
    if w * h:
        perims = clusters.device_perimeters()
        for idx, want in enumerate(ref.clusters):
            if want.area() == 0:
                assert perims[idx] == 0
                continue
            bits, bw, bh = R.to_image_with_hole(want, w, True)
            assert int(perims[idx]) == R.image_boundary_count(bits, bw, bh), idx
            assert clusters.device_neighbours(idx) == ref.neighbours_internal(want), idx
            for hole in (True, False):
                bits, bw, bh = R.to_image_with_hole(want, w, hole)
                im = clusters.device_to_image(idx, hole)
                assert (im.width, im.height) == (bw, bh) and bytes(im.pixels.astype(np.uint8)) == bytes(bits), (idx, hole)
    # row f1: to_image_with_hole(...).to_clusters(false) of every output cluster, batched on the device, against the independent
    # restatement of BinaryImage::to_clusters on the restated cluster image
    if w * h:
        for hole in (True, False):
            comps = clusters.device_components(hole)
            assert len(comps) == len(ref.clusters_output)
            for k, idx in enumerate(ref.clusters_output):
                bits, bw, bh = R.to_image_with_hole(ref.clusters[idx], w, hole)
                if bw * bh == 0:
                    assert comps[k] == []
                    continue
                want = R.to_clusters(np.array(list(bits), dtype=bool).reshape(bh, bw), False)
                assert len(comps[k]) == len(want["records"]), (k, idx, hole)
                for (rect, pts), wr in zip(comps[k], want["records"]):
                    o, n = int(wr["points_off"]), int(wr["points_len"])
                    assert rect == tuple(int(v) for v in wr["rect"]) and np.array_equal(pts, want["points"][o:o + n]), (k, idx, hole)
    if w * h:
        assert view.get_cluster_at(0) == ref.cluster_indices[0]
        assert view.get_cluster_at_point(api.PointI32(w - 1, h - 1)) == ref.cluster_indices[-1]
        assert view.get_pixel(0, 0).tuple() == tuple(int(v) for v in img[0, 0])
        assert view.get_pixel(-1, 0) is None and view.get_pixel(w, 0) is None and view.get_pixel_at_index(w * h) is None
        want_img = np.array(ref.to_color_image(), dtype=np.uint8).reshape(h, w, 4)
        assert np.array_equal(view.to_color_image().array(), want_img)
        canvas = api.ColorImage.new_w_h(w, h)
        for u in reversed(list(view.clusters_output)):
            view.get_cluster(int(u)).render_to_color_image(view, canvas)
        assert np.array_equal(canvas.array(), want_img)            # container.rs:104-116 spelled out on the host
    # Runner::run() gives the same object shape; take_image hands the pixels back
    again = api.Runner(cfg, api.ColorImage.from_array(img), ctx=ctx).run()
    assert [int(v) for v in again.view().cluster_indices] == ref.cluster_indices
    assert np.array_equal(again.take_image().array(), img)


def check_builder_contract(ctx):
    import pytest

    img = np.zeros((3, 3, 4), dtype=np.uint8)
    b = api.Runner(api.RunnerConfig(), api.ColorImage.from_array(img), ctx=ctx).builder()
    for name in ("same", "diff", "deepen", "hollow"):
        with pytest.raises(api.ClosureNotOffloadable):
            getattr(b, name)(lambda *a: True)
    assert b.diagonal(False).hierarchical(0).batch_size(100).key(api.Color()).keying_action(api.KeyingAction.Keep).run().output_len() >= 1
    with pytest.raises(AssertionError):                            # assert!(is_same_color_a < 8), runner.rs:78
        api.Runner(api.RunnerConfig(is_same_color_a=8), api.ColorImage.from_array(img), ctx=ctx).run()


def check_binary_api(ctx):
    # the reference's unit tests (clusters.rs:355-418) through the API mirror
    img = api.BinaryImage.from_string("*--\n-*-\n--*")
    cl = img.to_clusters(False, ctx=ctx)
    assert len(cl) == 3 and [(c.points[0].x, c.points[0].y) for c in cl] == [(0, 0), (1, 1), (2, 2)]
    assert (cl.clusters[2].rect.left, cl.clusters[2].rect.top, cl.clusters[2].rect.right, cl.clusters[2].rect.bottom) == (2, 2, 3, 3)
    assert cl.clusters[1].to_binary_image().get_pixel(0, 0)
    cl = img.to_clusters(True, ctx=ctx)
    assert len(cl) == 1 and [(p.x, p.y) for p in cl.clusters[0].points] == [(0, 0), (1, 1), (2, 2)]
    img = api.BinaryImage.new_w_h(4, 4)
    for y in (1, 2):
        for x in (1, 2):
            img.set_pixel(x, y, True)
    assert img.get_pixel(1, 1) and not img.get_pixel_safe(-1, 0) and not img.get_pixel_safe(4, 4)
    cl = img.to_clusters(False, ctx=ctx)
    assert len(cl) == 1 and (cl.rect.left, cl.rect.top, cl.rect.right, cl.rect.bottom) == (1, 1, 3, 3)
    assert api.BinaryImage.new_w_h(0, 0).to_clusters(False, ctx=ctx).clusters == []


def check_batch_queries(ctx, side=20):
    """device-resident queries / components on the handles of a BATCH (image index > 0: per-image offsets into the labels, the
    merge forest and the records) against the host mirror over the pools of the same handle."""
    import ctypes as C

    from parity import random_colour_image

    rng = np.random.default_rng(31)
    w, h = side, side - 3
    imgs = [random_colour_image(rng, w, h) for _ in range(3)]
    for cfg in (api.RunnerConfig(is_same_color_a=2, good_min_area=2, deepen_diff=8), api.RunnerConfig(hierarchical=0, is_same_color_a=2)):
        n = len(imgs)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        ws, hs = (C.c_uint32 * n)(*[w] * n), (C.c_uint32 * n)(*[h] * n)
        outs = (C.c_void_p * n)()
        ccfg = cfg.to_c()
        ctx.lib.check(ctx.lib.fn("runner_run_batch")(ctx.handle, C.byref(ccfg), ptrs, ws, hs, n, outs), ctx.handle)
        for k in range(n):
            cl = api.Clusters(ctx, C.c_void_p(outs[k]), imgs[k].reshape(-1), w, h)
            view = cl.view()
            perims = cl.device_perimeters()
            comps = cl.device_components(True)
            for pos, idx in enumerate(int(v) for v in view.clusters_output):
                c = view.get_cluster(idx)
                assert int(perims[idx]) == c.perimeter(view), (k, idx)
                assert cl.device_neighbours(idx) == c.neighbours(view), (k, idx)
                im = c.to_image(view)
                got = cl.device_to_image(idx, True)
                assert (got.width, got.height) == (im.width, im.height) and np.array_equal(got.pixels, im.pixels), (k, idx)
                want = R.to_clusters(im.pixels.reshape(im.height, im.width), False) if im.width * im.height else None
                assert len(comps[pos]) == (len(want["records"]) if want else 0), (k, idx)
                for (rect, pts), wr in zip(comps[pos], want["records"] if want else []):
                    o, ln = int(wr["points_off"]), int(wr["points_len"])
                    assert rect == tuple(int(v) for v in wr["rect"]) and np.array_equal(pts, want["points"][o:o + ln]), (k, idx)
