"""The reference's fake models (lib.rs:335-422), used by its own integration tests to pin the
pipeline conventions without real weights.  Test design reproduced, not code."""
import numpy as np


class FakeDetectionModel:
    """lib.rs:339-362: input [batch,1,200,100]; output = input + 0.5."""

    def input_shape(self):
        return ["batch", 1, 200, 100]

    def run(self, x):
        return (x + np.float32(0.5)).astype(np.float32)


class FakeRecognitionModel:
    """lib.rs:372-422: [N,1,64,W] -> max-pool W by 4 -> [W/4, N, 64]."""

    def input_shape(self):
        return ["batch", 1, 64, "seq"]

    def run(self, x):
        n, c, h, w = x.shape
        assert c == 1 and h == 64
        wb = w // 4
        pooled = x[:, 0, :, : wb * 4].reshape(n, h, wb, 4).max(axis=3)  # [N, H, W/4]
        return np.ascontiguousarray(np.transpose(pooled, (2, 0, 1))).astype(np.float32)


def gen_test_image(n_words: int) -> np.ndarray:
    """lib.rs:319-333: CHW f32 [3,100,200], black with n white 20x50 words on one line."""
    img = np.zeros((3, 100, 200), dtype=np.float32)
    for i in range(n_words):
        img[:, 30:50, i * 70: i * 70 + 50] = 1.0
    return img


def gen_rect_grid(top_left_yx, grid_shape, rect_size, gap_size):
    """test_util.rs:7-28 -> list of (top, left, bottom, right)."""
    y0, x0 = top_left_yx
    rows, cols = grid_shape
    rh, rw = rect_size
    gh, gw = gap_size
    out = []
    for r in range(rows):
        for c in range(cols):
            top = y0 + r * (rh + gh)
            left = x0 + c * (rw + gw)
            out.append((top, left, top + rh, left + rw))
    return out
