"""Drop-in for the reference's frame_enhancer module (frame_enhancer.py:1-238).

`ImageEnhancer` keeps the reference's constructor, attributes and methods;
every stage runs as a hand-written sm_100a kernel through libcvb200
(include/cvb200.h).  This module sits where the reference's selector loads
`src.cython.frame_enhancer_cython` (frame_enhancer.py:8-21,184-189); unlike
that selector it never falls back to a CPU class: no library / no GPU is an
error at construction time.
"""
import json
import os
import time

import numpy as np

from chessboard_vision_b200.engine import default_engine
from chessboard_vision_b200.hostapi import load_reference_module

USE_CYTHON = False          # the Cython twin is not used; kept because callers may read the flag
USE_B200 = True


class _ClaheB200:
    """Stands in for the cv2.CLAHE object the reference keeps in `self.clahe` (frame_enhancer.py:36)."""

    def __init__(self, engine, clip, tiles):
        self._e, self._clip, self._tiles = engine, float(clip), (int(tiles[0]), int(tiles[1]))

    def apply(self, plane):
        return self._e.clahe(np.ascontiguousarray(plane), self._clip, self._tiles)

    def getClipLimit(self):
        return self._clip

    def getTilesGridSize(self):
        return self._tiles

    def setClipLimit(self, v):
        self._clip = float(v)

    def setTilesGridSize(self, t):
        self._tiles = (int(t[0]), int(t[1]))


class ImageEnhancerB200:
    """B200 implementation of ImageEnhancerPython (frame_enhancer.py:23-181)."""

    def __init__(self, clahe_clip_limit=3.0, tile_grid_size=(8, 8), device=0):
        self._e = default_engine(device)
        self.clahe = _ClaheB200(self._e, clahe_clip_limit, tile_grid_size)
        self.sharpen_kernel = np.array([[-1, -1, -1], [-1, 9, -1], [-1, -1, -1]])
        self.profile = self.load_profile()

    def _params(self):
        return self._e.enhance_params(self.clahe.getClipLimit(), self.clahe.getTilesGridSize(), profile=self.profile)

    @staticmethod
    def _frame(frame):
        frame = np.asarray(frame)
        if frame.dtype != np.uint8 or frame.ndim != 3 or frame.shape[2] != 3:
            raise ValueError("expected a uint8 HxWx3 BGR frame, got %s %r" % (frame.dtype, frame.shape))
        return np.ascontiguousarray(frame)

    def load_profile(self):
        """frame_enhancer.py:46-54: color_profile.json in the working directory, if any."""
        try:
            if os.path.exists("color_profile.json"):
                with open("color_profile.json", "r") as f:
                    print("Loaded color profile")
                    return json.load(f)
        except Exception as e:
            print(f"Error loading profile: {e}")
        return {}

    def apply_color_profile(self, frame):
        """frame_enhancer.py:56-99: contrast/brightness, HSV hue / saturation / value adjustment (one pointwise
        kernel).  With no profile loaded (the configuration of the hot path) this is the identity."""
        if not self.profile:
            return frame
        return self._e.apply_color_profile(self._frame(frame), self.profile)

    def correct_lighting(self, frame):
        """frame_enhancer.py:101-120: BGR->LAB, CLAHE on L, LAB->BGR (fused kernel)."""
        return self._e.correct_lighting(self._frame(frame), self.clahe.getClipLimit(), self.clahe.getTilesGridSize())

    def reduce_noise(self, frame):
        """frame_enhancer.py:122-131: bilateral d=9, sigma 75/75."""
        return self._e.bilateral(self._frame(frame), 9, 75.0, 75.0)

    def sharpen(self, frame):
        """frame_enhancer.py:133-138: 3x3 sharpen."""
        return self._e.sharpen(self._frame(frame))

    def normalize_intensity(self, frame):
        """frame_enhancer.py:140-146: min-max normalize to 0..255."""
        return self._e.normalize(np.ascontiguousarray(frame))

    def prepare_analysis(self, frame):
        """frame_enhancer.py:148-159 -> (gray, binary)."""
        g, b = self._e.prepare_analysis(self._frame(frame))
        return g, b

    def process_pipeline(self, frame):
        """frame_enhancer.py:161-181: one call, four passes over the frame."""
        return self._e.process_pipeline(self._frame(frame), self._params())     # step 0 runs inside when a profile is set

    # not in the reference: both results of the demo loop (frame_enhancer.py:218-219) in one launch sequence
    def process_and_analyze(self, frame):
        enhanced, gray, binary, _ = self._e.enhance(self._frame(frame), self._params())
        return enhanced, gray, binary


def __getattr__(name):
    if name == "ImageEnhancerPython":       # the reference's CPU class, only if the checkout is importable
        ref = load_reference_module("frame_enhancer")
        if ref is not None:
            return ref.ImageEnhancerPython
    raise AttributeError(name)


ImageEnhancer = ImageEnhancerB200
print("[INFO] FrameEnhancer: B200 (sm_100a) backend")


def main():
    """frame_enhancer.py:192-235: webcam demo loop (needs a GUI build of OpenCV)."""
    import cv2
    cap = cv2.VideoCapture(0)
    if not cap.isOpened():
        print("Error: Could not open webcam.")
        return
    enhancer = ImageEnhancer()
    print("Starting Frame Enhancer... Press 'q' to quit.")
    prev = 0
    while True:
        ret, frame = cap.read()
        if not ret:
            print("Failed to grab frame.")
            break
        now = time.time()
        fps = 1 / (now - prev) if prev else 0
        prev = now
        enhanced, gray, binary = enhancer.process_and_analyze(frame)
        cv2.putText(frame, f"FPS: {int(fps)}", (10, 30), cv2.FONT_HERSHEY_SIMPLEX, 1, (0, 255, 0), 2)
        cv2.imshow('Original Feed', frame)
        cv2.imshow('Enhanced Feed', enhanced)
        cv2.imshow('Analysis (Otsu Binary)', binary)
        if cv2.waitKey(1) & 0xFF == ord('q'):
            break
    cap.release()
    cv2.destroyAllWindows()


if __name__ == "__main__":
    main()
