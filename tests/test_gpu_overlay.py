"""Board overlay on the GPU (cvb_overlay_dev through the C ABI): the reference's pictures (digests of the UNMODIFIED
GameSession._draw_interface), the cv2 call sequence of the oracle, random display lists against the NumPy interpreter,
batches, device-resident images and the argument errors."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import overlay_common as oc
from chessboard_vision_b200._lib import OverlayOp
from chessboard_vision_b200.overlay import BoardOverlay

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(HERE, "golden", "overlay.json")))


@pytest.mark.parametrize("name,size,state", oc.SCENARIOS, ids=[s[0] for s in oc.SCENARIOS])
def test_draw_interface_is_the_reference_picture(engine, name, size, state):
    from oracle import overlay as ov
    vis = oc.board_image(size)
    out = BoardOverlay(engine).draw_interface(vis, size, **state)
    assert out is vis                                            # drawn in place, as cv2 does
    assert oc.digest(vis) == GOLD[name]["sha256"]
    ref = ov.draw_interface_cv2(oc.board_image(size), size, **state)
    assert np.array_equal(vis, ref)


@pytest.mark.parametrize("seed,shape,n_ops", [(0, (97, 131), 40), (1, (480, 640), 200), (2, (33, 35), 64), (3, (800, 800), 1500),
                                              (4, (1, 1), 10), (5, (8, 2048), 120)])
def test_random_display_lists_against_the_interpreter(engine, seed, shape, n_ops):
    from oracle import overlay as ov
    rng = np.random.default_rng(seed)
    H, W = shape
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    dl = oc.random_display_list(rng, H, W, n_ops)
    ops, n, masks = dl.pack()
    assert n >= n_ops
    got = engine.overlay(img, ops, n, masks)
    assert np.array_equal(got, ov.apply_display_list(img, ops, n, masks))


def test_batch_and_device_resident(engine):
    from oracle import overlay as ov
    rng = np.random.default_rng(9)
    imgs = rng.integers(0, 256, (3, 120, 200, 3), dtype=np.uint8)
    dl = oc.random_display_list(rng, 120, 200, 50)
    ops, n, masks = dl.pack()
    got = engine.overlay(imgs, ops, n, masks)
    for i in range(3):
        assert np.array_equal(got[i], ov.apply_display_list(imgs[i], ops, n, masks))
    dev = engine.upload(imgs[0])
    assert BoardOverlay(engine).draw(dev, dl) is dev              # stays on the device, drawn in place
    assert np.array_equal(dev.get(), got[0])
    dev.free()
    assert np.array_equal(engine.overlay(imgs[0], ops, 0, b""), imgs[0])      # an empty list draws nothing


def test_argument_errors(engine):
    img = np.zeros((16, 16, 3), np.uint8)

    def one(**kw):
        f = dict(kind=0, x0=0, y0=0, x1=3, y1=3, alpha=1.0, beta=0.0, group=0, aux_ofs=0)
        f.update(kw)
        return OverlayOp(f["kind"], f["x0"], f["y0"], f["x1"], f["y1"], (C.c_uint8 * 4)(1, 2, 3, 0), f["alpha"], f["beta"], f["group"], f["aux_ofs"])
    for bad, masks in (([one(kind=7)], b""), ([one(kind=1, x1=-2)], b""), ([one(kind=2, x1=9, y1=2)], b"\x01\x02\x03"),
                       ([one(kind=2, x1=0, y1=2)], b"\x01"), ([one(alpha=float("nan"))], b""), ([one(x0=1 << 24)], b""),
                       ([one(group=1, alpha=0.5, beta=0.5), one(group=1, alpha=0.4, beta=0.6)], b""),
                       ([one(group=1), one(group=2), one(group=1)], b"")):
        arr = (OverlayOp * len(bad))(*bad)
        with pytest.raises(ValueError):                          # CVB_ERR_INVALID
            engine.overlay(img, arr, len(bad), masks)
    # corners in any order, as cv2.rectangle accepts them
    a = (OverlayOp * 1)(one(x0=9, y0=12, x1=2, y1=5))
    out = engine.overlay(img, a, 1, b"")
    assert out[5:13, 2:10].all() and out.sum() == 8 * 8 * 6
