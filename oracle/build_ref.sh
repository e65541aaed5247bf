#!/bin/bash
# TEST INFRASTRUCTURE ONLY. Builds the UNMODIFIED CheMPS2 reference library from the sources where
# they lie under /root/reference into oracle/_ref/ (git-ignored, travels to the GPU box).
# Missing system libraries are bridged by our own shims: env_shims/hdf5.h (in-memory HDF5 subset)
# and env_shims/blasfwd.c (dgemm_ & co -> OpenBLAS bundled with the scipy wheel).
# The reference's own cmake build is NOT run.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${CHEMPS2_REFERENCE:-/root/reference}
OUT="$HERE/_ref"
if [ ! -d "$REF/CheMPS2" ]; then
  echo "build_ref: $REF not present - using prebuilt $OUT if any"; exit 0
fi
SCIPYLIBS="$(python -c 'import scipy,os;print(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)),"scipy.libs"))')"
OPENBLAS="$(ls "$SCIPYLIBS"/libscipy_openblas*.so | head -1)"
mkdir -p "$OUT"
CXXFLAGS="-O2 -fopenmp -march=x86-64-v3 -fPIC -w -DH5_USE_110_API -DCHEMPS2_VERSION=\"1.8.12-oracle\" -I$HERE/../env_shims -I$REF/CheMPS2/include/chemps2"
if [ ! -f "$OUT/libblasfwd.so" ]; then
  gcc -O2 -fPIC -shared -o "$OUT/libblasfwd.so" "$HERE/../env_shims/blasfwd.c" "$OPENBLAS" -Wl,-rpath,"$SCIPYLIBS"
fi
if [ ! -f "$OUT/libchemps2.so" ] || [ "$HERE/../env_shims/hdf5.h" -nt "$OUT/libchemps2.so" ]; then
  mkdir -p "$OUT/obj"
  ls "$REF"/CheMPS2/*.cpp | grep -v executable.cpp | \
    xargs -P "$(nproc)" -I{} sh -c "g++ $CXXFLAGS -c {} -o $OUT/obj/\$(basename {} .cpp).o"
  g++ -shared -fopenmp -o "$OUT/libchemps2.so.new" "$OUT"/obj/*.o -L"$OUT" -lblasfwd -Wl,-rpath,'$ORIGIN' -Wl,-rpath,"$SCIPYLIBS"
  mv -f "$OUT/libchemps2.so.new" "$OUT/libchemps2.so"
  rm -rf "$OUT/obj" "$OUT/ref_driver" "$OUT/chemps2"
fi
# the oracle driver (our code: dumps fixtures / times the reference hot path through its own classes)
if [ -f "$HERE/ref_driver.cpp" ] && { [ ! -f "$OUT/ref_driver" ] || [ "$HERE/ref_driver.cpp" -nt "$OUT/ref_driver" ]; }; then
  g++ $CXXFLAGS -o "$OUT/ref_driver.new" "$HERE/ref_driver.cpp" -L"$OUT" -lchemps2 -lblasfwd -Wl,-rpath,'$ORIGIN' -Wl,-rpath,"$SCIPYLIBS"
  mv -f "$OUT/ref_driver.new" "$OUT/ref_driver"   # atomic swap: a running ref_driver keeps its old file
fi
# the reference's own command-line binary (expected outputs of the drop-in tests come from it)
if [ ! -f "$OUT/chemps2" ]; then
  g++ $CXXFLAGS -o "$OUT/chemps2" "$REF/CheMPS2/executable.cpp" -L"$OUT" -lchemps2 -lblasfwd -Wl,-rpath,'$ORIGIN' -Wl,-rpath,"$SCIPYLIBS"
fi
# input data of the whole-sweep comparison in bench.py (--sweep-ref): the reference's own N2/cc-pVDZ FCIDUMP (config 2); git-ignored like the rest of _ref
cp -f "$REF/tests/matrixelements/N2.CCPVDZ.FCIDUMP" "$OUT/" 2>/dev/null || true
echo "build_ref: ok -> $OUT"
