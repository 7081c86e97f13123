"""Device-resident per-square state behind dict-like host views.

The reference keeps `PieceDetector.reference_squares`, `ChangeDetector.means`
and `.variances` as plain dicts of per-square ndarrays that callers may read
or assign (piece_detector.py:48,95-97; change_detector.py:29-30,91-92).  Here
the arrays live on the GPU as planes shaped like the board image (each square
owns its rectangle); `PlaneDict` is the dict-like window onto one plane:
reads download the plane lazily, writes are staged and flushed before the
next kernel launch.

Reading `d[key]` hands out ONE host array per key until the device rewrites the
plane: in-place edits of it (`d[key][...] = x`, `d[key] *= k`, which the
reference's plain dicts allow) are noticed at the next launch and uploaded.
After a kernel that rewrites the plane the entry is a new array, exactly as the
reference rebinds `self.means[pos] = new_mean` (change_detector.py:91-92).
"""
from collections.abc import MutableMapping

import numpy as np

from . import _lib
from .hostapi import pack_squares


class PlaneDict(MutableMapping):
    def __init__(self, runner, plane, dtype, flag_bit):
        self._r, self._plane, self._dtype, self._bit = runner, plane, np.dtype(dtype), flag_bit
        self._present = set()       # keys that hold a value (device or staged)
        self._staged = {}           # key -> ndarray assigned by the host, not yet on the device
        self._removed = False       # a key was deleted since the last flush
        self._hostonly = {}         # values whose key / shape is not part of the current board layout
        self._host = None           # downloaded copy of the device plane
        self._live = {}             # key -> (array handed to the caller, pristine copy) since the last device write

    # -- Mapping protocol --
    def __getitem__(self, key):
        if key in self._staged:
            return self._staged[key]
        if key in self._hostonly:
            return self._hostonly[key]
        if key not in self._present:
            raise KeyError(key)
        rect = self._r.layout.get(key)
        if rect is None:
            raise KeyError(key)
        if self._host is None:
            self._host = self._r.state.get(0, self._plane)
        if key not in self._live:
            x, y, w, h = rect
            arr = self._host[y:y + h, x:x + w].copy()
            self._live[key] = (arr, arr.copy())
        return self._live[key][0]

    def __setitem__(self, key, value):
        self._staged[key] = np.array(value, dtype=self._dtype)
        self._live.pop(key, None)
        self._hostonly.pop(key, None)
        self._present.add(key)

    def __delitem__(self, key):
        if key not in self._present:
            raise KeyError(key)
        self._present.discard(key)
        self._staged.pop(key, None)
        self._live.pop(key, None)
        self._hostonly.pop(key, None)
        self._removed = True

    def __iter__(self):
        return iter([k for k in self._r.key_order() if k in self._present] +
                    [k for k in list(self._staged) + list(self._hostonly) if k not in self._r.layout])

    def __len__(self):
        return len(self._present)

    def __contains__(self, key):
        return key in self._present

    def clear(self):
        if self._present:
            self._removed = True
        self._present.clear()
        self._staged.clear()
        self._live.clear()
        self._hostonly.clear()

    # -- used by the runner --
    def dirty(self):
        # arrays handed out by __getitem__ and edited in place since become staged writes
        for key, (arr, orig) in list(self._live.items()):
            if not np.array_equal(arr, orig):
                self._staged[key] = arr
                del self._live[key]
        return bool(self._staged) or self._removed

    def device_changed(self, keys_now_present=()):
        """The kernel rewrote (part of) the plane."""
        self._host = None
        self._live.clear()
        self._present.update(keys_now_present)

    def snapshot(self):
        """All current values as host arrays (used when the board layout changes)."""
        return {k: self[k] for k in list(self._present)}


class SquareRunner:
    """Owns one cvb_state (one stream slot) and keeps it consistent with the PlaneDicts
    and with the geometry of the square dictionaries passed in."""

    def __init__(self, engine):
        self.e = engine
        self.state = None
        self.layout = {}            # key -> (x, y, w, h) of the current board layout
        self._order = []
        self._shape = None
        self.dicts = []             # PlaneDicts bound to this runner

    def bind(self, plane, dtype, flag_bit):
        d = PlaneDict(self, plane, dtype, flag_bit)
        self.dicts.append(d)
        return d

    def key_order(self):
        return self._order

    def _rebuild(self, shape, layout, order):
        saved = [d.snapshot() if self.state is not None or d._staged else {} for d in self.dicts]
        if self.state is not None:
            self.state.free()
        self.state = self.e.new_state(1, shape[0], shape[1])
        self.layout, self._order, self._shape = dict(layout), list(order), shape
        for d, vals in zip(self.dicts, saved):
            d._host = None
            d._live = {}
            d._present = set(vals.keys())
            d._staged = dict(vals)      # re-upload through the flush path
            d._hostonly = {}
            d._removed = True

    def _flush(self):
        if not any(d.dirty() for d in self.dicts):
            return
        flags = self.state.get(0, _lib.PLANE_FLAGS)
        for d in self.dicts:
            if not d.dirty():
                continue
            plane = self.state.get(0, d._plane)
            for k, (x, y, w, h) in self.layout.items():
                if k not in d._present:
                    flags[y:y + h, x:x + w] &= np.uint8(~d._bit & 0xff)
            for k, v in list(d._staged.items()):
                rect = self.layout.get(k)
                if rect is None or v.shape != (rect[3], rect[2]):
                    d._hostonly[k] = d._staged.pop(k)   # not part of this layout: host-only entry
                    continue
                x, y, w, h = rect
                plane[y:y + h, x:x + w] = v
                flags[y:y + h, x:x + w] |= np.uint8(d._bit)
                del d._staged[k]
            d._removed = False
            self.state.set(0, d._plane, plane)
            d._host = plane
        self.state.set(0, _lib.PLANE_FLAGS, flags)

    def run(self, squares, params, select_keys=None, want_stats=True):
        """Launch the square kernel on a dict of squares -> ({key: stats record}, keys)."""
        board, rects, keys = pack_squares(squares)
        if board is None:
            return {}, []
        layout = dict(zip(keys, rects))
        shape = board.shape[:2]
        if self.state is None or shape != self._shape or any(self.layout.get(k, r) != r for k, r in layout.items()):
            self._rebuild(shape, layout, keys)
        else:
            for k in keys:                       # same geometry, possibly a different subset of squares
                if k not in self.layout:
                    self.layout[k] = layout[k]
                    self._order.append(k)
        self._flush()
        select = None
        if select_keys is not None:
            sel = set(select_keys)
            select = np.array([1 if k in sel else 0 for k in keys], np.uint8)
        stats = self.e.squares(board, rects, params, self.state, 0, select, want_stats)
        return ({k: stats[0, i] for i, k in enumerate(keys)} if want_stats else {}), keys

    def current_plane(self):
        """gray+blur squares of the last launch (PLANE_PD_CUR)."""
        return self.state.get(0, _lib.PLANE_PD_CUR)

    def hough(self, keys, params):
        """cv2.HoughCircles on the gray+blur squares of the last launch, for `keys` only -> {key: record}."""
        order = [k for k in self._order if k in self.layout]
        want = set(keys)
        select = np.array([1 if k in want else 0 for k in order], np.uint8)
        if not select.any():
            return {}
        out = self.e.hough_state(self.state, [self.layout[k] for k in order], params, 0, 1, select)
        return {k: out[0, i] for i, k in enumerate(order) if k in want}
