#!/bin/bash
# One command on the first box that has cargo AND a B200: produce reference-held golden vectors from the real
# qp-plonky2 =1.5.5, then run the parity tests that consume them (oracle on CPU, CUDA path on the GPU).
#   scripts/pin_on_rust_box.sh            # from the repository root
# Afterwards commit tests/golden/reference_vectors.json: from then on `pytest -m "not gpu"` checks the oracle against
# the reference's own numbers and `pytest -m gpu` checks the kernels, and the "parity unpinned" note in oracle/ and
# DESIGN.md can go for every entry the file covers.  Next step: the [patch.crates-io] fork + rust/qpzk-sys
# and the reference's own tests (wormhole/tests: prover_tests.rs:15-20, verifier_tests.rs:39-66, circuit_data_tests.rs:69-128).
set -euo pipefail
cd "$(dirname "$0")/.."
command -v cargo >/dev/null || { echo "no cargo on this box: parity stays unpinned"; exit 2; }
( cd rust/golden-dump && cargo run --release -- ../../tests/golden/reference_vectors.json )
python -c "import __graft_entry__ as g; g.build()"
python -m pytest tests/test_reference_vectors.py -q -m "not gpu"
if python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)"; then
  python -m pytest tests/test_reference_vectors.py -q -m gpu
fi
