#!/bin/bash
# smoke() with the current library and with a library built from another commit (gpurun_tmp/libmaua_base.so)
cd "${GRAFT_REPO_ROOT:-.}"
run() { timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-250; }
echo "== current"; run
cp maua_style_b200/libmaua_b200.so /tmp/new.so
cp gpurun_tmp/libmaua_base.so maua_style_b200/libmaua_b200.so
echo "== base library"; MAUA_DEV_ALLOW_MISSING=1 run
cp /tmp/new.so maua_style_b200/libmaua_b200.so
