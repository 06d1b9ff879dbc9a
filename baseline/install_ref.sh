#!/bin/bash
# Install the UNMODIFIED reference (gprMax v3.1.7, /root/reference) into baseline/_ref (git-ignored, shipped to the GPU box
# by gpurun) so that its own front end -- input parsing, geometry / material / PML build, run_model, the .out writer --
# can run next to a B200 with only the time loop replaced.  Runs in the build container only.
#   * the source tree is read-only and setup.py cythonises in place -> install from a copy under /tmp;
#   * h5py / colorama / terminaltables / matplotlib are not in the wheelhouse -> --no-deps (baseline/standins.py provides
#     inert stand-ins for the three that are imported on the run path);
#   * setup.py hard-codes -march=native -> baseline/cc_portable.sh pins x86-64-v3 (the binaries run on another host);
#   * user_models/*.in (input FILES, BASELINE.json configs 1, 3, 4) are not package data -> copied next to the package.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${GPRMAX_REFERENCE:-/root/reference}"
TMP="$(mktemp -d /tmp/gprmax_ref_XXXX)"
cp -r "$REF"/. "$TMP"/
rm -rf "$HERE/_ref"
( cd "$TMP" && CC="$HERE/cc_portable.sh" LDSHARED="$HERE/cc_portable.sh -shared" \
  python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP" )
mkdir -p "$HERE/_ref/user_models"
cp "$REF"/user_models/*.in "$HERE/_ref/user_models/"
rm -rf "$TMP"
echo "installed: $(ls "$HERE/_ref")"
