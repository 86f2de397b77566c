#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds the UNMODIFIED reference (chrchang/plink-ng 2.0) from the
# sources where they lie under /root/reference into oracle/_ref/ (git-ignored, travels to the
# GPU box with gpurun).  Our own recipe: g++/gcc on the reference's source files directly; the
# reference's Makefiles are not run.  Source list = CSRC+ZCSRC+ZSSRC+CCSRC of
# /root/reference/2.0/Makefile.src:8-111; flags follow build_dynamic/Makefile:41,128-135
# (-ffp-contract=off, AVX2, -O2).
#
#   oracle/_ref/plink2         NOLAPACK build (KING, GRM naive loop, --indep-pairwise)
#   oracle/_ref/plink2_lapack  LAPACK build against the OpenBLAS bundled in the venv (PCA, BLAS GRM)
#
# Nothing in the product links or runs these; only tests/ and bench.py's cpu_baseline /
# --impl reference legs execute them.
set -euo pipefail
REF=${REF:-/root/reference/2.0}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
JOBS=${JOBS:-$(nproc)}
if [ ! -d "$REF" ]; then
  echo "build_ref.sh: $REF absent (GPU box?) - using prebuilt $OUT if present"; exit 0
fi
mkdir -p "$OUT/obj" "$OUT/obj_lapack" "$OUT/stub"

CSRC="include/SFMT.c libdeflate/lib/adler32.c libdeflate/lib/crc32.c libdeflate/lib/deflate_compress.c
 libdeflate/lib/deflate_decompress.c libdeflate/lib/gzip_compress.c libdeflate/lib/gzip_decompress.c
 libdeflate/lib/utils.c libdeflate/lib/zlib_compress.c libdeflate/lib/zlib_decompress.c
 libdeflate/lib/arm/arm_cpu_features.c libdeflate/lib/x86/x86_cpu_features.c"
ZCSRC=$(cd "$REF" && ls zstd/lib/common/*.c zstd/lib/compress/*.c zstd/lib/decompress/*.c)
ZSSRC="zstd/lib/decompress/huf_decompress_amd64.S"
CCSRC=$(cd "$REF" && ls include/plink2_*.cc include/pgenlib_misc.cc include/pgenlib_read.cc include/pgenlib_write.cc plink2*.cc | grep -v plink2_cpu.cc)

BASE="-DZSTD_MULTITHREAD -ffp-contract=off -mavx2 -mbmi -mbmi2 -mfma -mlzcnt"
INC="-I$REF/libdeflate -I$REF/libdeflate/common -I$REF/zstd/lib -I$REF/zstd/lib/common -I$REF/simde"

objname() { echo "$1" | sed 's#[/.]#_#g'; }

compile_list=$OUT/obj/compile.txt; : > "$compile_list"
for f in $CSRC; do echo "gcc -O2 -std=gnu99 $BASE $INC -w -c $REF/$f -o $OUT/obj/$(objname $f).o" >> "$compile_list"; done
for f in $ZCSRC $ZSSRC; do echo "gcc -O2 -std=gnu99 $BASE $INC -w -c $REF/$f -o $OUT/obj/$(objname $f).o" >> "$compile_list"; done
for f in $CCSRC; do
  echo "g++ -std=c++17 -O2 $BASE $INC -DNOLAPACK -include math.h -w -c $REF/$f -o $OUT/obj/$(objname $f).o" >> "$compile_list"
done
echo "[build_ref] compiling $(wc -l < "$compile_list") objects (NOLAPACK) with $JOBS jobs"
# skip objects that are already newer than their source
xargs -P "$JOBS" -I{} bash -c '{}' < "$compile_list"
g++ $OUT/obj/*.o -o $OUT/plink2 -lm -lpthread -lz
echo "[build_ref] built $OUT/plink2"

# ---- LAPACK variant: only headers are missing on this image; write stand-ins for exactly the
# symbols plink2 uses (cblas_* / Fortran LAPACK), link the venv's OpenBLAS.
OBLIB=$(python3 - <<'EOF'
import glob, sys, sysconfig
p = glob.glob(sysconfig.get_paths()["purelib"] + "/opencv_python_headless.libs/libopenblasp-*.so")
print(p[0] if p else "")
EOF
)
if [ -z "$OBLIB" ]; then echo "[build_ref] no bundled OpenBLAS; skipping plink2_lapack"; exit 0; fi
cat > $OUT/stub/cblas.h <<'EOF'
#ifndef STUB_CBLAS_H
#define STUB_CBLAS_H
#ifdef __cplusplus
extern "C" {
#endif
enum CBLAS_ORDER {CblasRowMajor=101, CblasColMajor=102};
enum CBLAS_TRANSPOSE {CblasNoTrans=111, CblasTrans=112, CblasConjTrans=113};
enum CBLAS_UPLO {CblasUpper=121, CblasLower=122};
typedef enum CBLAS_ORDER CBLAS_LAYOUT;
double cblas_ddot(const int n, const double* x, const int incx, const double* y, const int incy);
float cblas_sdot(const int n, const float* x, const int incx, const float* y, const int incy);
double cblas_dsdot(const int n, const float* x, const int incx, const float* y, const int incy);
void cblas_dgemm(const enum CBLAS_ORDER, const enum CBLAS_TRANSPOSE, const enum CBLAS_TRANSPOSE, const int, const int, const int, const double, const double*, const int, const double*, const int, const double, double*, const int);
void cblas_sgemm(const enum CBLAS_ORDER, const enum CBLAS_TRANSPOSE, const enum CBLAS_TRANSPOSE, const int, const int, const int, const float, const float*, const int, const float*, const int, const float, float*, const int);
void cblas_dgemv(const enum CBLAS_ORDER, const enum CBLAS_TRANSPOSE, const int, const int, const double, const double*, const int, const double*, const int, const double, double*, const int);
void cblas_sgemv(const enum CBLAS_ORDER, const enum CBLAS_TRANSPOSE, const int, const int, const float, const float*, const int, const float*, const int, const float, float*, const int);
void cblas_dsyrk(const enum CBLAS_ORDER, const enum CBLAS_UPLO, const enum CBLAS_TRANSPOSE, const int, const int, const double, const double*, const int, const double, double*, const int);
void cblas_ssyrk(const enum CBLAS_ORDER, const enum CBLAS_UPLO, const enum CBLAS_TRANSPOSE, const int, const int, const float, const float*, const int, const float, float*, const int);
void openblas_set_num_threads(int);
#ifdef __cplusplus
}
#endif
#endif
EOF
cat > $OUT/stub/lapacke.h <<'EOF'
#ifndef STUB_LAPACKE_H
#define STUB_LAPACKE_H
#include <stdint.h>
#define lapack_int int32_t
#define LAPACK_dgecon dgecon_
#define LAPACK_dgesvd dgesvd_
#define LAPACK_dgetrf dgetrf_
#define LAPACK_dgetri dgetri_
#define LAPACK_dlange dlange_
#define LAPACK_dlansy dlansy_
#define LAPACK_dpocon dpocon_
#define LAPACK_dpotrf dpotrf_
#define LAPACK_dpotri dpotri_
#define LAPACK_dpotrs dpotrs_
#define LAPACK_dsyevr dsyevr_
#define LAPACK_sgetrf sgetrf_
#define LAPACK_sgetri sgetri_
#define LAPACK_spotrf spotrf_
#define LAPACK_spotri spotri_
#define LAPACK_ssyevr ssyevr_
#ifdef __cplusplus
extern "C" {
#endif
void dgecon_(const char* norm, const lapack_int* n, const double* a, const lapack_int* lda, const double* anorm, double* rcond, double* work, lapack_int* iwork, lapack_int* info);
void dgesvd_(const char* jobu, const char* jobvt, const lapack_int* m, const lapack_int* n, double* a, const lapack_int* lda, double* s, double* u, const lapack_int* ldu, double* vt, const lapack_int* ldvt, double* work, const lapack_int* lwork, lapack_int* info);
void dgetrf_(const lapack_int* m, const lapack_int* n, double* a, const lapack_int* lda, lapack_int* ipiv, lapack_int* info);
void dgetri_(const lapack_int* n, double* a, const lapack_int* lda, const lapack_int* ipiv, double* work, const lapack_int* lwork, lapack_int* info);
double dlange_(const char* norm, const lapack_int* m, const lapack_int* n, const double* a, const lapack_int* lda, double* work);
double dlansy_(const char* norm, const char* uplo, const lapack_int* n, const double* a, const lapack_int* lda, double* work);
void dpocon_(const char* uplo, const lapack_int* n, const double* a, const lapack_int* lda, const double* anorm, double* rcond, double* work, lapack_int* iwork, lapack_int* info);
void dpotrf_(const char* uplo, const lapack_int* n, double* a, const lapack_int* lda, lapack_int* info);
void dpotri_(const char* uplo, const lapack_int* n, double* a, const lapack_int* lda, lapack_int* info);
void dpotrs_(const char* uplo, const lapack_int* n, const lapack_int* nrhs, const double* a, const lapack_int* lda, double* b, const lapack_int* ldb, lapack_int* info);
void dsyevr_(const char* jobz, const char* range, const char* uplo, const lapack_int* n, double* a, const lapack_int* lda, const double* vl, const double* vu, const lapack_int* il, const lapack_int* iu, const double* abstol, lapack_int* m, double* w, double* z, const lapack_int* ldz, lapack_int* isuppz, double* work, const lapack_int* lwork, lapack_int* iwork, const lapack_int* liwork, lapack_int* info);
void sgetrf_(const lapack_int* m, const lapack_int* n, float* a, const lapack_int* lda, lapack_int* ipiv, lapack_int* info);
void sgetri_(const lapack_int* n, float* a, const lapack_int* lda, const lapack_int* ipiv, float* work, const lapack_int* lwork, lapack_int* info);
void spotrf_(const char* uplo, const lapack_int* n, float* a, const lapack_int* lda, lapack_int* info);
void spotri_(const char* uplo, const lapack_int* n, float* a, const lapack_int* lda, lapack_int* info);
void ssyevr_(const char* jobz, const char* range, const char* uplo, const lapack_int* n, float* a, const lapack_int* lda, const float* vl, const float* vu, const lapack_int* il, const lapack_int* iu, const float* abstol, lapack_int* m, float* w, float* z, const lapack_int* ldz, lapack_int* isuppz, float* work, const lapack_int* lwork, lapack_int* iwork, const lapack_int* liwork, lapack_int* info);
#ifdef __cplusplus
}
#endif
#endif
EOF
compile_list=$OUT/obj_lapack/compile.txt; : > "$compile_list"
for f in $CCSRC; do
  echo "g++ -std=c++17 -O2 $BASE $INC -DUSE_OPENBLAS -I$OUT/stub -w -c $REF/$f -o $OUT/obj_lapack/$(objname $f).o" >> "$compile_list"
done
echo "[build_ref] compiling $(wc -l < "$compile_list") objects (LAPACK)"
xargs -P "$JOBS" -I{} bash -c '{}' < "$compile_list"
COBJ=$(ls $OUT/obj/*.o | grep -v -E 'obj/(include_p|plink2)' )
OBDIR=$(dirname "$OBLIB")
g++ $OUT/obj_lapack/*.o $COBJ -o $OUT/plink2_lapack -L"$OBDIR" -l:$(basename "$OBLIB") -Wl,--disable-new-dtags -Wl,-rpath,"$OBDIR" -lm -lpthread -lz
echo "[build_ref] built $OUT/plink2_lapack"
