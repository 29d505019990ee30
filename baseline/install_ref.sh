#!/usr/bin/env bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored, shipped to the GPU box by gpurun).
# 1) the sanctioned offline pip install; 2) if the build backend (poetry-core) is missing offline -- it is,
#    see DESIGN.md -- fall back to copying the reference's single source file verbatim.
set -u
cd "$(dirname "$0")/.."
mkdir -p baseline/_ref
if python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse \
      --target baseline/_ref /root/reference >/tmp/ref_install.log 2>&1; then
  echo "pip install of the reference succeeded"
elif python -m pip install --no-deps --no-index --no-build-isolation --find-links /opt/wheelhouse \
      --target baseline/_ref /root/reference >>/tmp/ref_install.log 2>&1; then
  echo "pip install --no-deps of the reference succeeded"
else
  echo "pip install failed (poetry-core build backend not available offline; the project declares no packages):"
  tail -n 3 /tmp/ref_install.log
fi
# The reference is a loose script (pyproject declares no packages), so pip would not ship model.py anyway.
if [ ! -f baseline/_ref/model.py ]; then
  cp /root/reference/model.py baseline/_ref/model.py
  echo "copied /root/reference/model.py verbatim to baseline/_ref/model.py"
fi
cmp /root/reference/model.py baseline/_ref/model.py && echo "baseline/_ref/model.py is byte-identical to the reference"
