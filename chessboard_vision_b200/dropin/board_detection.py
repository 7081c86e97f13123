"""Drop-in for the reference's board_detection module: `warp_image`
(board_detection.py:61-71) runs on the B200 (perspective matrix on the host in
f64, resampling kernel on the device) and `reorder` (board_detection.py:49-58)
is host index logic.  `find_chessboard_corners` (board_detection.py:4-27) runs its
image part (gray, blur, Canny, dilation) on the B200 and the contour logic (a sequential
border-following algorithm on the resulting mask) with host OpenCV, as the reference does.
Every other name of the reference module (grid drawing: GUI code outside the hot path) is
forwarded to the reference's own board_detection.py when it is on sys.path.
"""
import numpy as np

from chessboard_vision_b200.engine import default_engine
from chessboard_vision_b200.hostapi import load_reference_module


def reorder(myPoints):
    """Order four corner points TL, TR, BL, BR by coordinate sum / difference (board_detection.py:49-58)."""
    pts = np.asarray(myPoints).reshape((4, 2))
    out = np.zeros((4, 1, 2), np.int32)
    s = pts.sum(1)
    d = np.diff(pts, axis=1)
    out[0] = pts[np.argmin(s)]
    out[3] = pts[np.argmax(s)]
    out[1] = pts[np.argmin(d)]
    out[2] = pts[np.argmax(d)]
    return out


def rectContour(contours):
    """Contours with area > 100000 that simplify to four corners, largest first (board_detection.py:30-39)."""
    import cv2
    rect = []
    for c in contours:
        if cv2.contourArea(c) > 100000:
            approx = cv2.approxPolyDP(c, 0.02 * cv2.arcLength(c, True), True)
            if len(approx) == 4:
                rect.append(c)
    return sorted(rect, key=cv2.contourArea, reverse=True)


def getCornerPoints(contour):
    """board_detection.py:42-45"""
    import cv2
    return cv2.approxPolyDP(contour, 0.02 * cv2.arcLength(contour, True), True)


def find_chessboard_corners(img, debug=False):
    """board_detection.py:4-27 -> the four corners (TL, TR, BL, BR) of the largest four-sided contour,
    or an empty array.  The mask comes from the GPU; `debug` (cv2.imshow in the reference) is ignored."""
    import cv2
    img = np.asarray(img)
    if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
        raise ValueError("expected a uint8 HxWx3 BGR frame, got %s %r" % (img.dtype, img.shape))
    mask = default_engine().contour_mask(np.ascontiguousarray(img))
    contours, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    rect = rectContour(contours)
    if len(rect) == 0:
        return np.array([])
    biggest = getCornerPoints(rect[0])
    if biggest.size != 0:
        return reorder(biggest)
    return np.array([])


def warp_image(img, points, display_size=(1280, 720), margin=100):
    """-> (img_warped, matrix, board_size) like board_detection.warp_image."""
    e = default_engine()
    board_size = min(display_size) - margin
    img = np.asarray(img)
    if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
        raise ValueError("expected a uint8 HxWx3 BGR frame, got %s %r" % (img.dtype, img.shape))
    pts1 = np.float32(points).reshape(4, 2)
    pts2 = np.float32([[0, 0], [board_size, 0], [0, board_size], [board_size, board_size]])
    matrix = e.get_perspective_transform(pts1, pts2)
    warped = e.warp(np.ascontiguousarray(img), matrix, board_size)
    return warped, matrix, board_size


def __getattr__(name):
    ref = load_reference_module("board_detection")
    if ref is not None and hasattr(ref, name):
        return getattr(ref, name)
    raise AttributeError("board_detection.%s is outside the B200 hot path and the reference module was not found" % name)
