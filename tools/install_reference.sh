#!/bin/bash
# Puts an UNMODIFIED copy of the reference's Python tree for the train / render path under baseline/_ref/ (git-ignored, not
# gpurun-ignored: it travels to the GPU box, where /root/reference does not exist) so that
#     python -m b200gs.launcher --reference baseline/_ref train_4DGS.py ...
# can be exercised on hardware (tests/test_launcher_gpu.py, bench.py's `launcher_path` block).  Only files the two scripts import
# are copied (no stage-1 code, no third-party trees, no fonts); nothing is edited.  Run where /root/reference exists.
set -e
REF=${REF_ROOT:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
DST=$ROOT/baseline/_ref
[ -d "$REF/scene" ] || { echo "no reference tree at $REF"; exit 0; }
rm -rf "$DST"; mkdir -p "$DST/utils"
cp "$REF/train_4DGS.py" "$REF/render_4DGS.py" "$DST/"
cp -r "$REF/scene" "$REF/gaussian_renderer" "$REF/arguments" "$REF/test_trajectory" "$DST/"
cp "$REF"/utils/*.py "$DST/utils/"
find "$DST" -name __pycache__ -prune -exec rm -rf {} +
( cd "$REF" && sha256sum train_4DGS.py render_4DGS.py gaussian_renderer/__init__.py scene/gaussian_model.py scene/deformation.py scene/hexplane.py ) > "$DST/SHA256SUMS"
echo "installed $(find "$DST" -type f | wc -l) files into $DST"
