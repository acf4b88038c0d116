#!/usr/bin/env bash
# Stages the UNMODIFIED reference (torchode 1.0.1, pure Python) under baseline/_ref/ so that it can
# travel to the GPU box with the gpurun snapshot (baseline/_ref/ is git-ignored, not gpurun-ignored):
#   baseline/_ref/torchode/      verbatim copy of /root/reference/torchode
#   baseline/_ref/torchtyping/   annotation-only stub (the reference imports torchtyping purely for
#                                type annotations: typing.py:4,50-87, runge_kutta.py:5,15-28)
#   baseline/_ref/ref_tests/     verbatim copy of /root/reference/tests (the reference's own suite)
# `pip install --target baseline/_ref /root/reference` is not possible here: the build backend
# (flit_core) is neither installed nor in /opt/wheelhouse.  Used by: bench.py --impl reference,
# bench.py's reference_cuda leg, tests/test_gpu_vs_reference.py, tests/test_reference_suite.py.
set -euo pipefail
SRC="${1:-/root/reference}"
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
DST="$ROOT/baseline/_ref"
[ -d "$SRC/torchode" ] || { echo "no reference at $SRC" >&2; exit 1; }
rm -rf "$DST/torchode" "$DST/torchtyping" "$DST/ref_tests"
mkdir -p "$DST/torchtyping"
cp -r "$SRC/torchode" "$DST/torchode"
cp -r "$SRC/tests" "$DST/ref_tests"
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
cat > "$DST/torchtyping/__init__.py" <<'PY'
"""Annotation-only stand-in for torchtyping (not installed in this image)."""


class _Meta(type):
    def __getitem__(cls, item):
        return cls


class TensorType(metaclass=_Meta):
    pass


is_float = object()
PY
( cd "$SRC" && find torchode tests -type f -name '*.py' | sort | xargs sha256sum ) > "$DST/MANIFEST.sha256"
echo "staged $(find "$DST/torchode" -name '*.py' | wc -l) reference files under $DST"
