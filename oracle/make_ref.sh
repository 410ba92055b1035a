#!/bin/bash
# TEST / BENCH INFRASTRUCTURE ONLY.  Stages the reference's OWN pure-Python implementation of the loss path where it
# can travel to the GPU box: the files oracle/ref_import.py imports are copied, unmodified, from
# /root/reference into the git-ignored oracle/_ref/ (never into history; gpurun ships the directory).  bench.py's
# reference arm and cpu_baseline then time the reference itself ("kind": "reference"); without oracle/_ref they fall
# back to the restatement oracle/port.py ("kind": "port").  Run in the build container: bash oracle/make_ref.sh
set -e
SRC="${SDE_REFERENCE_ROOT:-/root/reference}/detectron2"
DST="$(cd "$(dirname "$0")" && pwd)/_ref/detectron2"
if [ ! -d "$SRC" ]; then echo "make_ref: $SRC not found (nothing staged)"; exit 0; fi
FILES="geometry/camera.py geometry/resampler.py geometry/pose_utils.py
modeling/losses/__init__.py modeling/losses/losses.py modeling/losses/motion_loss.py modeling/losses/photometric_loss.py
modeling/losses/smoothness_loss.py modeling/losses/ssim_loss.py
modeling/meta_arch/MonoDepth2.py modeling/meta_arch/MotionLearning.py utils/memory.py"
rm -rf "$DST"
for f in $FILES; do
  mkdir -p "$DST/$(dirname "$f")"
  cp "$SRC/$f" "$DST/$f"
done
( cd "$DST" && sha256sum $FILES ) > "$DST/../SHA256SUMS"
echo "make_ref: staged $(echo $FILES | wc -w) reference files under $DST"
