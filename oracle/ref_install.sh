#!/usr/bin/env bash
# Installs the UNMODIFIED reference into baseline/_ref/ (git-ignored; travels to the GPU box) on top of libpmpc_b200.so:
#
#   baseline/_ref/pmpc/                     the reference's Python package (pmpc/*.py)
#   baseline/_ref/pmpcjl.<abi>.so           the reference's own pybind11 shim, PMPC.jl/pmpcjl/module.cpp, compiled
#                                           from where it lies under /root/reference, linked against libpmpc_b200.so
#                                           instead of libPMPC.so + libjulia.so (the library exports c_lqp_solve,
#                                           c_lcone_solve and the no-op init_julia the shim calls at import)
#   baseline/_ref/PMPC.jl/pmpcjl/lib/       the directory pmpc/import_pmpcjl.py:12-38 preloads lib*.so from
#
# so that `PYTHONPATH=baseline/_ref python -c "import pmpc; pmpc.solve(...)"` runs the reference's own scp_solve ->
# static_backend.aff_solve -> pmpcjl.lqp_solve -> c_lqp_solve on the GPU (tests/test_reference_dropin.py).
#
# `pip install --target baseline/_ref /root/reference` is tried first, as the task contract asks.  It cannot succeed in
# this image: the reference's setup.py build hook imports the `julia` package and installs PMPC.jl (setup.py:44-86), and
# neither Julia nor pyjulia exists here (no network).  The fallback places the same pure-Python package files pip would
# have placed.  Needs /root/reference, i.e. runs in the build container only.
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${PMPC_REFERENCE:-/root/reference}"
DST="$ROOT/baseline/_ref"
[ -d "$REF" ] || { echo "ref_install: $REF not present (GPU box): nothing to do"; exit 0; }
mkdir -p "$DST"
TMP="$(mktemp -d)"
cp -r "$REF" "$TMP/ref"
if python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target "$DST" "$TMP/ref" \
      > "$DST/pip_install.log" 2>&1; then
  echo "ref_install: pip install succeeded"
else
  echo "ref_install: pip install failed ($(grep -m1 -o "ModuleNotFoundError.*" "$DST/pip_install.log" || echo "see baseline/_ref/pip_install.log")); placing the pure-Python package"
  rm -rf "$DST/pmpc"
  cp -r "$REF/pmpc" "$DST/pmpc" && chmod -R u+w "$DST/pmpc"
fi
rm -rf "$TMP"
# the reference's pybind shim, unmodified, against libpmpc_b200.so
EXT="$(python3-config --extension-suffix 2>/dev/null || python -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
g++ -O2 -shared -fPIC -std=c++11 -w $(python -m pybind11 --includes) \
    -I"$REF/PMPC.jl/pmpcjl/include" "$REF/PMPC.jl/pmpcjl/module.cpp" -o "$DST/pmpcjl$EXT" \
    -L"$ROOT/pmpc_b200" -lpmpc_b200 -Wl,-rpath,'$ORIGIN/../../pmpc_b200' || exit 1
mkdir -p "$DST/PMPC.jl/pmpcjl/lib"
ln -sf ../../../../../pmpc_b200/libpmpc_b200.so "$DST/PMPC.jl/pmpcjl/lib/libpmpc_b200.so"
echo "ref_install: baseline/_ref ready ($(ls "$DST" | tr '\n' ' '))"
