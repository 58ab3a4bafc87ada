#!/usr/bin/env bash
# oracle/make_ref.sh -- stage the UNMODIFIED reference for the GPU box.
#
# Test / baseline infrastructure, not product code.  Copies the files of the hot path, its
# callers (the scripts north_star says must "run unchanged") and the shipped checkpoints from the read-only
# reference tree into oracle/_ref/ with the reference's own directory layout.  oracle/_ref/ is git-ignored (no
# reference source enters the history) but not gpurun-ignored, so it travels to the box like the built .so, where
#   * bench.py --impl reference and bench.py's cpu_baseline time the reference's own classes (kind "reference"),
#   * tests/test_gpu_ref_scripts.py runs the reference's scripts byte-for-byte with optimalmodulationds_b200/dropin
#     first on PYTHONPATH.
# Nothing is edited: the two import shims the reference needs under torch 2.11 without matplotlib live outside the
# copied tree (oracle/shims/, oracle/ref_harness.py).  Re-run after the reference changes; idempotent.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${DSMPPI_REFERENCE_SRC:-/root/reference}"
DST="$HERE/_ref"
PS="python_scripts"
if [ ! -d "$SRC/$PS/ds_mppi/functions" ]; then
  echo "make_ref.sh: no reference tree at $SRC (nothing staged)" >&2
  exit 3
fi
rm -rf "$DST"
mkdir -p "$DST/$PS/ds_mppi/functions" "$DST/$PS/ds_mppi/scripts" "$DST/$PS/mlp_learn/sdf" "$DST/$PS/mlp_learn/models"
# the hot path and the modules its star-imports pull in
for f in MPPI policy cost LinDS fk_num plots MPPI_toy cost_toy SEDS fk_sym_gen zmq_utils; do
  cp "$SRC/$PS/ds_mppi/functions/$f.py" "$DST/$PS/ds_mppi/functions/"
done
cp "$SRC/$PS/mlp_learn/sdf/"*.py "$DST/$PS/mlp_learn/sdf/"
# callers: single-process demos, the dense-field plot script, the Franka planner / integrator processes + their config
for f in standalonePlanar2d standalonePlanar7d standaloneToy2d standalonePlanar2d_policyPlots; do
  cp "$SRC/$PS/ds_mppi/scripts/$f.py" "$DST/$PS/ds_mppi/scripts/"
done
cp "$SRC/$PS/ds_mppi/scripts/toy_policy.pt" "$DST/$PS/ds_mppi/scripts/"
for f in frankaPlanner frankaIntegrator obstacleStreamer; do
  cp "$SRC/$PS/ds_mppi/$f.py" "$DST/$PS/ds_mppi/"
done
cp "$SRC/$PS/ds_mppi/config.yaml" "$DST/$PS/ds_mppi/"
mkdir -p "$DST/$PS/ds_mppi/experiment_logs"      # frankaIntegrator.py:100 appends its run log here
# SEDS mixtures (content/ds/*.mat, referenced from frankaIntegrator.py:70-71; a few KB each)
if [ -d "$SRC/$PS/ds_mppi/content/ds" ]; then
  mkdir -p "$DST/$PS/ds_mppi/content"
  cp -r "$SRC/$PS/ds_mppi/content/ds" "$DST/$PS/ds_mppi/content/"
fi
# shipped checkpoints of the three robots + the toy net
for f in 2dof_sdf_256x5_mesh 7dof_sdf_256x5_mesh franka_collision_model 2dof_sdf_256x5_toy; do
  cp "$SRC/$PS/mlp_learn/models/$f.pt" "$DST/$PS/mlp_learn/models/"
done
( cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo "unknown" ) > "$DST/REFERENCE_COMMIT"
( cd "$DST" && find . -type f ! -name MANIFEST.sha256 | sort | xargs sha256sum ) > "$DST/MANIFEST.sha256"
echo "make_ref.sh: staged $(find "$DST" -type f | wc -l) files, $(du -sh "$DST" | cut -f1) in $DST"
