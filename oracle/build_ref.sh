#!/bin/bash
# ORACLE tooling: stage the UNMODIFIED reference modules of the hot path under oracle/_ref/ so that they travel to the
# GPU box with the snapshot (oracle/_ref/ is git-ignored: nothing of the reference enters the history; it is NOT
# gpurun-ignored).  The reference is pure Python: there is nothing to compile, "building" the reference
# arm is copying the files byte for byte and recording their checksums.
#   bench.py --impl reference   drives these modules on the host cores (cpu_baseline.kind "reference")
#   tests/test_gpu_dropin.py    runs them ON TOP of msfno_b200 (INTEGRATION.md levels 1 and 2)
# Only tests/, __graft_entry__ and bench.py's CPU legs ever import from here (test_product_does_not_import_oracle).
set -e
SRC=${MSFNO_REFERENCE_ROOT:-/root/reference}
DST="$(cd "$(dirname "$0")" && pwd)/_ref"
if [ ! -d "$SRC/MSFNO/Models/sfno" ]; then
  echo "reference tree not mounted at $SRC: keeping the staged copy in $DST (if any)"; exit 0
fi
FILES="MSFNO/__init__.py MSFNO/utils.py MSFNO/Models/__init__.py MSFNO/Models/losses.py
MSFNO/Models/sfno/__init__.py MSFNO/Models/sfno/sfnonet.py MSFNO/Models/sfno/layers.py
MSFNO/Models/sfno/contractions.py MSFNO/Models/sfno/activations.py
MSFNO/Models/gcn/gcn.py MSFNO/Models/gcn/layers.py MSFNO/Models/vit/vit.py MSFNO/Models/mae/maenet.py"
rm -rf "$DST"
for f in $FILES; do
  mkdir -p "$DST/$(dirname $f)"
  cp "$SRC/$f" "$DST/$f"
done
( cd "$SRC" && sha256sum $FILES ) > "$DST/SHA256SUMS"
( cd "$DST" && sha256sum -c SHA256SUMS --quiet )
echo "staged $(echo $FILES | wc -w) unmodified reference files in $DST"
