#!/bin/bash
# Compile the reference's own sources IN PLACE (from /root/reference, nothing is copied) into oracle/_ref/libmrcpp_ref.so,
# against the Eigen stand-in under oracle/eigen_shim (Eigen 3.4.0 is not available in this image). Outputs only under
# oracle/_ref/. TEST INFRASTRUCTURE: an independent check of oracle/oracle.cpp and the CPU baseline of bench.py.
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
[ -d "$REF/src" ] || { echo "no reference tree at $REF"; exit 0; }
mkdir -p $OUT/include/MRCPP $OUT/obj
for d in core functions operators treebuilders trees utils; do ln -sfn $REF/src/$d $OUT/include/MRCPP/$d; done
for f in $REF/api/*; do b=$(basename $f); [ "$b" = CMakeLists.txt ] || ln -sf $f $OUT/include/MRCPP/$b; done
cat > $OUT/include/MRCPP/config.h <<EOC
#pragma once
namespace mrcpp {
inline constexpr auto mwfilters_source_dir() noexcept { return "$REF/share/mwfilters"; }
inline constexpr auto mwfilters_install_dir() noexcept { return "$REF/share/mwfilters"; }
} // namespace mrcpp
#define BLAS_H ""
EOC
cat > $OUT/include/MRCPP/version.h <<EOC
#pragma once
namespace mrcpp {
inline constexpr auto program_version() noexcept { return "reference sources compiled in place"; }
inline constexpr auto git_commit_hash() noexcept { return "unknown"; }
inline constexpr auto git_commit_author() noexcept { return "unknown"; }
inline constexpr auto git_commit_date() noexcept { return "unknown"; }
inline constexpr auto git_branch() noexcept { return "unknown"; }
} // namespace mrcpp
EOC
CXXFLAGS="-std=c++17 -O3 -march=x86-64-v3 -fopenmp -fPIC -DMRCPP_HAS_OMP -w -I$HERE/eigen_shim -I$OUT/include -I$REF/src -I$REF/api"
compile_one() {
  src=$1
  obj=$OUT/obj/$(echo $src | sed 's#^\./##; s#/#_#g; s#\.cpp$#.o#')
  if [ ! -f $obj ] || [ $REF/src/$src -nt $obj ] || [ $HERE/eigen_shim/Eigen/Core -nt $obj ]; then
    g++ $CXXFLAGS -c $REF/src/$src -o $obj 2> $obj.err || { echo "FAILED $src (see $obj.err)"; rm -f $obj; }
  fi
}
export -f compile_one
export OUT REF HERE CXXFLAGS
(cd $REF/src && find . -name '*.cpp' | sort) | xargs -P $(nproc) -I{} bash -c 'compile_one {}'
n_src=$(cd $REF/src && find . -name '*.cpp' | wc -l)
n_obj=$(ls $OUT/obj/*.o 2>/dev/null | wc -l)
[ "$n_src" = "$n_obj" ] || { echo "reference build incomplete: $n_obj of $n_src objects"; exit 1; }
g++ -shared -fopenmp -o $OUT/libmrcpp_ref.so $OUT/obj/*.o
# the C interface the tests and bench.py's reference arm bind to (oracle/ref_driver.cpp: our code, the reference's public API)
g++ $CXXFLAGS -shared $HERE/ref_driver.cpp -o $OUT/libref_driver.so -L$OUT -lmrcpp_ref -Wl,-rpath,'$ORIGIN'
# filter / cross-correlation tables for MWFILTERS_DIR: the same numerical data the product ships (mrcpp_b200/data/mwtables.bin),
# unpacked into the file names the reference reads (MWFilter.cpp:200-251, CrossCorrelation.cpp:99-124), so that the library
# also runs where /root/reference does not exist (the GPU box)
python3 - "$HERE/../mrcpp_b200/data/mwtables.bin" "$OUT/mwfilters" <<'EOP'
import os, struct, sys
src, dst = sys.argv[1], sys.argv[2]
os.makedirs(dst, exist_ok=True)
names = {0: "I_H0_%d", 1: "I_G0_%d", 2: "I_c_left_%d", 3: "I_c_right_%d"}
# tabulated derivative matrices: text files read sequentially from K = 2 up to the order asked for (PHCalculator.cpp:47-79)
text = {4: "I_ph_deriv_1.txt", 5: "I_ph_deriv_2.txt", 6: "I_b-spline-deriv1.txt", 7: "I_b-spline-deriv2.txt", 8: "I_b-spline-deriv3.txt"}
blocks = {kind: {} for kind in text}
with open(src, "rb") as f:
    assert f.read(4) == b"MRXT"
    (n,) = struct.unpack("<i", f.read(4))
    for _ in range(n):
        kind, k, cnt = struct.unpack("<iii", f.read(12))
        data = f.read(8 * cnt)
        if kind in text:
            blocks[kind][k + 1] = struct.unpack("<%dd" % cnt, data)
            continue
        with open(os.path.join(dst, names[kind] % k), "wb") as g:
            g.write(data)
for kind, name in text.items():
    with open(os.path.join(dst, name), "w") as g:
        K = 2
        while K in blocks[kind]:
            g.write("%d\n" % K)
            v = blocks[kind][K]
            for i in range(3 * K):
                g.write(" ".join("%.17e" % x for x in v[i * K:(i + 1) * K]) + " \n")
            K += 1
EOP
# the reference-side binding of INTEGRATION.md as a real program (tests/cpp/ref_binding.cpp: the reference's own FunctionTree /
# ConvolutionOperator handed to the C ABI of libmrcpp_b200.so): one binary for the GPU box, one with the device entry points
# served by the CPU oracle (tests/cpp/oracle_backend.cpp) for the CPU suite. Needs the product library (built first).
PROD=$HERE/../mrcpp_b200/lib
if [ -f $PROD/libmrcpp_b200.so ]; then
  BFLAGS="$CXXFLAGS -I$HERE/../include -L$OUT -lmrcpp_ref -L$PROD -lmrcpp_b200 -ldl -Wl,-rpath,\$ORIGIN -Wl,-rpath,\$ORIGIN/../../mrcpp_b200/lib"
  if [ ! -f $OUT/ref_binding ] || [ $HERE/../tests/cpp/ref_binding.cpp -nt $OUT/ref_binding ] || [ $HERE/../include/mrcpp_b200.h -nt $OUT/ref_binding ]; then
    g++ $HERE/../tests/cpp/ref_binding.cpp $BFLAGS -o $OUT/ref_binding
    g++ -rdynamic $HERE/../tests/cpp/ref_binding.cpp $HERE/../tests/cpp/oracle_backend.cpp $BFLAGS -o $OUT/ref_binding_cpu
  fi
fi
rm -rf $OUT/include  # build-time symlinks into the reference tree: nothing that points outside the repo may travel
echo "built $OUT/libmrcpp_ref.so $OUT/libref_driver.so $OUT/mwfilters"
