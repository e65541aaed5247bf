#!/bin/bash
# Builds the DROP-IN libchemps2.so.3: the UNMODIFIED reference sources (compiled where they lie under /root/reference, public and private
# headers byte-identical) with the hot-path member functions replaced by dropin/chemps2_b200_shim.cpp over the C ABI of
# chemps2_b200/libchemps2_b200.so.  The replacement happens at the object level: the reference's own definitions of
#   Heff::SolveDAVIDSON / makeHeff / fillHeffDiag  and  DMRG::updateMovingRight / updateMovingLeft
# are demoted to weak symbols (objcopy --weaken-symbol), the shim's strong definitions win at link time.  No reference source is copied
# or edited.  Also builds the reference's `chemps2` binary and its own tests (tests/testN.cpp.in with the data path substituted) against
# the drop-in library.  Outputs only into dropin/_build/ (git-ignored; travels to the GPU box).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF=${CHEMPS2_REFERENCE:-/root/reference}
OUT="$HERE/_build"
if [ ! -d "$REF/CheMPS2" ]; then echo "build_dropin: $REF not present - using prebuilt $OUT if any"; exit 0; fi
if [ ! -f "$ROOT/chemps2_b200/libchemps2_b200.so" ]; then echo "build_dropin: build chemps2_b200/libchemps2_b200.so first (make)"; exit 1; fi
SCIPYLIBS="$(python -c 'import scipy,os;print(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)),"scipy.libs"))')"
OPENBLAS="$(ls "$SCIPYLIBS"/libscipy_openblas*.so | head -1)"
mkdir -p "$OUT/obj" "$OUT/tests/matrixelements"
CXXFLAGS="-O2 -fopenmp -march=x86-64-v3 -fPIC -w -DH5_USE_110_API -DCHEMPS2_VERSION=\"1.8.12-b200\" -I$ROOT/env_shims -I$REF/CheMPS2/include/chemps2"
[ -f "$OUT/libblasfwd.so" ] || gcc -O2 -fPIC -shared -o "$OUT/libblasfwd.so" "$ROOT/env_shims/blasfwd.c" "$OPENBLAS" -Wl,-rpath,"$SCIPYLIBS"
# 1. the unmodified reference, one object per source
if [ ! -f "$OUT/obj/.done" ]; then
  ls "$REF"/CheMPS2/*.cpp | grep -v executable.cpp | xargs -P "$(nproc)" -I{} sh -c "g++ $CXXFLAGS -c {} -o $OUT/obj/\$(basename {} .cpp).o"
  touch "$OUT/obj/.done"
fi
# 2. demote the reference's definitions of the five hot-path members to weak symbols
weaken() {  # object, c++filt pattern
  for sym in $(nm "$1" | awk '$2 == "T" {print $3}'); do
    if c++filt "$sym" | grep -q "^$2("; then objcopy --weaken-symbol="$sym" "$1"; echo "   weak: $(c++filt "$sym" | cut -c1-80)"; fi
  done
}
weaken "$OUT/obj/Heff.o" "CheMPS2::Heff::SolveDAVIDSON"
weaken "$OUT/obj/Heff.o" "CheMPS2::Heff::makeHeff"
weaken "$OUT/obj/Heff.o" "CheMPS2::Heff::fillHeffDiag"
weaken "$OUT/obj/DMRGoperators.o" "CheMPS2::DMRG::updateMovingRight"
weaken "$OUT/obj/DMRGoperators.o" "CheMPS2::DMRG::updateMovingLeft"
# 3. the shim and the library (SONAME 3, CMakeLists.txt:12)
g++ $CXXFLAGS -Wall -I"$ROOT/include" -c "$HERE/chemps2_b200_shim.cpp" -o "$OUT/chemps2_b200_shim.o"
g++ -shared -fopenmp -Wl,-soname,libchemps2.so.3 -o "$OUT/libchemps2.so.3" "$OUT"/obj/*.o "$OUT/chemps2_b200_shim.o" \
    -L"$OUT" -lblasfwd -L"$ROOT/chemps2_b200" -lchemps2_b200 -Wl,-rpath,'$ORIGIN' -Wl,-rpath,'$ORIGIN/../../chemps2_b200' -Wl,-rpath,"$SCIPYLIBS"
ln -sf libchemps2.so.3 "$OUT/libchemps2.so"
LINK="-L$OUT -lchemps2 -lblasfwd -Wl,-rpath,\$ORIGIN -Wl,-rpath,\$ORIGIN/../../chemps2_b200 -Wl,-rpath,$SCIPYLIBS -Wl,-rpath-link,$ROOT/chemps2_b200"
# 4. the reference's binary and its own tests, linked against the drop-in
g++ $CXXFLAGS -o "$OUT/chemps2" "$REF/CheMPS2/executable.cpp" $LINK
cp -f "$REF"/tests/matrixelements/*.FCIDUMP "$OUT/tests/matrixelements/"
for f in test2.input test14.input; do   # the inputs of the binary, FCIDUMP path pointing at the copied data files
  sed "s#/path/to/#$OUT/tests/matrixelements/#" "$REF/tests/$f" > "$OUT/tests/$f"
done
for n in 1 2 3 4 5 6 7 8 9 10 11 12 13; do
  [ -f "$REF/tests/test$n.cpp.in" ] || continue
  sed "s#\${CMAKE_SOURCE_DIR}#$OUT#g" "$REF/tests/test$n.cpp.in" > "$OUT/obj/test$n.cpp"
  g++ $CXXFLAGS -o "$OUT/test$n" "$OUT/obj/test$n.cpp" $LINK &
done
wait
rm -f "$OUT"/obj/test*.cpp
echo "build_dropin: ok -> $OUT"
