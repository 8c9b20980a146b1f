#!/usr/bin/env bash
# Stage the UNMODIFIED reference (anibali/dsnt-pose2d) under baseline/_ref with the one offline install the contract names:
#
#   python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference>
#
# baseline/_ref is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so it travels to the GPU
# box, where /root/reference does not exist.  It is used only as (a) the reference arm / CPU baseline of bench.py
# (`kind: "reference"`), (b) the eager-torch GPU baseline of bench.py, and (c) the subject of tests/test_gpu_ref_models.py,
# which drives the reference's own ResNet / hourglass model classes with this library's head attached.
# --no-deps: the reference's only declared dependency is torch (already here); its import-time extras (torchdata.mpii,
# tele, torchnet ...) are absent from the image and are stubbed by the tests where dsnt.model needs them.
# The install runs from a copy under /tmp because setuptools writes an egg-info into the source tree and
# /root/reference is read-only.
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
REF="${1:-/root/reference}"
if [ ! -f "$REF/setup.py" ]; then
  echo "install_ref.sh: no reference checkout at $REF (nothing to do)" >&2
  exit 0
fi
TMP="$(mktemp -d /tmp/dsnt_ref_XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF" "$TMP/src"
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
  --target "$ROOT/baseline/_ref" "$TMP/src"
find "$ROOT/baseline/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
test -f "$ROOT/baseline/_ref/dsnt/nn.py"
echo "install_ref.sh: reference installed under baseline/_ref ($(ls "$ROOT/baseline/_ref" | tr '\n' ' '))"
