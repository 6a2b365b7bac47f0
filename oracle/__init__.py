"""CPU oracle for the ocrs hot path.  TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

This package is a CPU restatement (numpy / torch-CPU / small pure-Python loops) of the
reference pipeline `preprocess -> detect -> layout -> recognise -> CTC` of
robertknight/ocrs @ 4bccf6b (v0.12.2).  Every function cites the reference
`file:line` it follows (paths relative to the reference checkout).

Who may import it: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`, and only as the *checker*.
Nothing under `ocrs_b200/` imports it; the product path fails loudly when the CUDA
library is missing rather than falling back to this code.

PARITY STATUS (see DESIGN.md "Oracle"):
  * pipeline stages whose source is in the reference tree (`ocrs/src/*.rs`) are
    restated line by line and pinned by the reference's own known-answer tests,
    ported in `tests/test_oracle_goldens.py`.
  * everything that lives in the un-vendored crates `rten`, `rten-imageproc`,
    `rten-tensor` 0.24.0 (Cargo.lock:682-786) -- the two networks, bilinear resize,
    contour tracing, RDP, min-area-rect, polygon fill, CTC decode -- is restated from the
    published algorithms (ONNX operator spec, Suzuki-Abe 1985, Ramer-Douglas-Peucker,
    exhaustive hull-edge min-area rectangle, even-odd scanline fill) and anchored on the
    reference's call sites and fake-model goldens.  There is no rten source, binary or
    weight file in this environment: for those functions **parity is unpinned** beyond
    what the reference's own goldens cover.
"""

BLACK_VALUE = -0.5  # preprocess.rs:128
