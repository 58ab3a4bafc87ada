#!/bin/bash
# Builds the instrumented library (-DDSMPPI_TCX_PROF) as libdsmppi_b200_prof.so beside the product library, then
# rebuilds the product library; tools/tcx_prof.py picks the instrumented one up when it exists.
set -e
cd "$(dirname "$0")/.."
DSMPPI_EXTRA_NVCC_FLAGS=-DDSMPPI_TCX_PROF python -c "from optimalmodulationds_b200 import build as b; b.build(force=True)"
cp optimalmodulationds_b200/libdsmppi_b200.so optimalmodulationds_b200/libdsmppi_b200_prof.so
python -c "from optimalmodulationds_b200 import build as b; print(b.build(force=True))"
