#!/usr/bin/env bash
# Installs the UNMODIFIED reference (KinWaiCheuk/nnAudio v0.3.3) into baseline/_ref (git-ignored,
# NOT gpurun-ignored: it travels to the GPU box) so that `bench.py --impl reference` and the
# `reference_gpu` leg time the reference's own code.  Needs /root/reference (build container only).
# The source tree is read-only and setup.py writes build/ + egg-info, so install from a /tmp copy;
# --no-deps because the wheelhouse resolver cannot see the already-installed numpy/torch.
set -euo pipefail
root="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
src="${NNAUDIO_REF_SRC:-/root/reference/Installation}"
[ -d "$src" ] || { echo "no reference at $src (GPU box: uses the prebuilt baseline/_ref)"; exit 0; }
tmp="$(mktemp -d)"
cp -r "$src" "$tmp/ref"
rm -rf "$root/baseline/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps \
    --find-links /opt/wheelhouse --target "$root/baseline/_ref" "$tmp/ref"
rm -rf "$tmp"
echo "installed reference into $root/baseline/_ref"
