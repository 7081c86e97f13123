"""Board overlay, CPU side: the oracle (cv2 call sequence) against the digests of the unmodified
GameSession._draw_interface, the display-list builder + glyph cache against that oracle through the NumPy interpreter
of the list, and the arithmetic facts the kernel relies on (addWeighted, circle spans, text translation)."""
import json
import os
import sys

import cv2
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import overlay_common as oc
from oracle import overlay as ov
from chessboard_vision_b200.overlay import BoardOverlay, DisplayList

GOLD = json.load(open(os.path.join(HERE, "golden", "overlay.json")))


@pytest.mark.parametrize("name,size,state", oc.SCENARIOS, ids=[s[0] for s in oc.SCENARIOS])
def test_oracle_matches_unmodified_reference(name, size, state):
    vis = oc.board_image(size)
    assert oc.digest(vis) == GOLD[name]["input_sha256"]
    ov.draw_interface_cv2(vis, size, **state)
    assert oc.digest(vis) == GOLD[name]["sha256"]
    if name == "small_board":
        ref = np.load(os.path.join(HERE, "golden", "overlay_small.npz"))["small_board"]
        assert np.array_equal(vis, ref)


@pytest.mark.parametrize("name,size,state", oc.SCENARIOS, ids=[s[0] for s in oc.SCENARIOS])
def test_display_list_reproduces_the_drawing(name, size, state):
    """host logic: the list BoardOverlay builds, interpreted with OpenCV's arithmetic, is the reference's picture"""
    dl = BoardOverlay.display_list(size, **state)
    ops, n, masks = dl.pack()
    got = ov.apply_display_list(oc.board_image(size), ops, n, masks)
    assert oc.digest(got) == GOLD[name]["sha256"]


def test_add_weighted_formula_all_byte_pairs():
    v = np.arange(256, dtype=np.uint8)
    a, b = np.meshgrid(v, v, indexing="ij")
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    for alpha, beta in ((0.3, 0.7), (0.5, 0.5), (0.4, 0.6), (0.6, 0.4), (0.25, 0.8), (1.0, 1.0)):
        assert np.array_equal(cv2.addWeighted(a, alpha, b, beta, 0), ov.add_weighted_u8(a, alpha, b, beta)), (alpha, beta)
    # odd lengths take OpenCV's scalar tail: same arithmetic
    rng = np.random.default_rng(0)
    for n in (1, 3, 7, 17, 33):
        x, y = rng.integers(0, 256, (1, n), dtype=np.uint8), rng.integers(0, 256, (1, n), dtype=np.uint8)
        assert np.array_equal(cv2.addWeighted(x, 0.3, y, 0.7, 0), ov.add_weighted_u8(x, 0.3, y, 0.7))


def test_blend_of_a_copy_is_the_identity():
    """addWeighted(vis.copy(), a, vis, b) leaves pixels outside the drawn shapes unchanged: the kernel skips them"""
    v = np.arange(256, dtype=np.uint8).reshape(1, -1)
    for alpha, beta in ((0.3, 0.7), (0.5, 0.5), (0.4, 0.6), (0.6, 0.4)):
        assert np.array_equal(cv2.addWeighted(v.copy(), alpha, v, beta, 0), v)
        assert np.array_equal(ov.add_weighted_u8(v, alpha, v, beta), v)


def test_circle_spans_match_cv2():
    for r in range(0, 72):
        img = np.zeros((160, 160), np.uint8)
        cv2.circle(img, (80, 80), r, 255, -1)
        mine = np.zeros_like(img)
        for dy, hw in enumerate(ov.circle_half_widths(r)):
            mine[80 - dy, 80 - hw:80 + hw + 1] = 255
            mine[80 + dy, 80 - hw:80 + hw + 1] = 255
        assert np.array_equal(img, mine), r


def test_text_stamps_translate_and_clip_like_puttext():
    rng = np.random.default_rng(1)
    font = cv2.FONT_HERSHEY_SIMPLEX
    for text, scale, thick in (("Q", 1.2, 4), ("q", 1.2, 2), ("jogada em andamento", 1.0, 3), ("FPS: 59.9", 0.6, 2), ("Turno: Pretas", 0.6, 2)):
        for _ in range(6):
            org = (int(rng.integers(-40, 230)), int(rng.integers(-10, 240)))
            ref = rng.integers(0, 256, (200, 220, 3), dtype=np.uint8)
            base = ref.copy()
            cv2.putText(ref, text, org, font, scale, (7, 200, 90), thick)
            dl = DisplayList()
            dl.put_text(text, org, font, scale, (7, 200, 90), thick)
            ops, n, masks = dl.pack()
            assert np.array_equal(ov.apply_display_list(base, ops, n, masks), ref), (text, org)


def test_display_list_structure():
    dl = BoardOverlay.display_list(800, pieces=oc.START, white_to_move=True, last_move=((4, 1), (4, 3)), radar=[(1, 1)])
    ops, n, masks = dl.pack()
    kinds = [ops[i].kind for i in range(n)]
    assert kinds[:18] == [0] * 18 and kinds.count(1) == 1
    assert n == 18 + 2 + 1 + 2 * 32 + 2                       # lines, last move, radar, 32 pieces x 2 passes, 2 status texts
    groups = [ops[i].group for i in range(n) if ops[i].group]
    assert groups == [1, 1]                                   # both squares of the last move share one overlay copy
    with pytest.raises(ValueError):
        DisplayList().line((0, 0), (5, 7), (1, 2, 3))
