#!/usr/bin/env bash
# Install the UNMODIFIED reference (BloodAxe/pytorch-toolbelt, mounted read-only at /root/reference) into
# baseline/_ref so that it travels to the GPU box with the repo snapshot (baseline/_ref is git-ignored, not
# gpurun-ignored). Used only by `bench.py --impl reference` and by the parity tests that compare against the real
# reference when it is present. Nothing under pytorch_toolbelt_b200/ ever imports it.
#
# The source tree is read-only and setuptools wants to write egg-info next to setup.py, so the build runs from a
# scratch copy under /tmp; dependency resolution is skipped (--no-deps): torch / numpy / cv2 / scipy are already in
# the image and the offline wheelhouse holds no torch wheel to "resolve" against.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${1:-/root/reference}"
if [ ! -d "$SRC" ]; then
  echo "install_ref: $SRC not found (fine on the GPU box: baseline/_ref ships prebuilt)"; exit 0
fi
TMP="$(mktemp -d /tmp/tb_ref_XXXXXX)"
cp -r "$SRC" "$TMP/src"
rm -rf "$HERE/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
  --target "$HERE/_ref" "$TMP/src"
rm -rf "$TMP"
echo "install_ref: installed $(ls "$HERE/_ref" | tr '\n' ' ')"
