#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (oracle/).  Builds the two vendored third-party libraries the reference's
# Functional kernels need -- serial MFEM and axom core+slic -- from the sources where they lie under
# /root/reference, into oracle/_ref/ (git-ignored).  Nothing under /root/reference is modified and no
# reference source is copied into the repository.  Recipe follows SURVEY.md Appendix A.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${SERAC_REFERENCE:-/root/reference}"
R="$HERE/_ref"
J="${JOBS:-8}"
mkdir -p "$R"
if [ ! -d "$REF/mfem" ]; then echo "no reference at $REF: skipping (prebuilt oracle/_ref is used as-is)"; exit 0; fi

if [ ! -f "$R/mfem-install/lib/libmfem.a" ]; then
  cmake -S "$REF/mfem" -B "$R/mfem-build" -DCMAKE_BUILD_TYPE=Release -DMFEM_USE_MPI=NO \
        -DMFEM_ENABLE_TESTING=OFF -DMFEM_ENABLE_EXAMPLES=OFF -DMFEM_ENABLE_MINIAPPS=OFF \
        -DCMAKE_POSITION_INDEPENDENT_CODE=ON -DCMAKE_INSTALL_PREFIX="$R/mfem-install" > "$R/mfem-cmake.log" 2>&1
  make -C "$R/mfem-build" -j"$J" install > "$R/mfem-make.log" 2>&1
fi
if [ ! -f "$R/axom-install/lib/libaxom_core.a" ]; then
  cmake -S "$REF/axom/src" -B "$R/axom-build" -DCMAKE_BUILD_TYPE=Release -DENABLE_MPI=OFF -DAXOM_ENABLE_MPI=OFF \
        -DAXOM_ENABLE_ALL_COMPONENTS=OFF -DAXOM_ENABLE_SLIC=ON -DAXOM_ENABLE_TESTS=OFF -DAXOM_ENABLE_EXAMPLES=OFF \
        -DAXOM_ENABLE_DOCS=OFF -DAXOM_ENABLE_TOOLS=OFF -DENABLE_TESTS=OFF -DENABLE_DOCS=OFF -DENABLE_EXAMPLES=OFF \
        -DENABLE_FORTRAN=OFF -DENABLE_OPENMP=OFF -DCMAKE_POSITION_INDEPENDENT_CODE=ON \
        -DCMAKE_INSTALL_PREFIX="$R/axom-install" > "$R/axom-cmake.log" 2>&1
  make -C "$R/axom-build" -j"$J" install > "$R/axom-make.log" 2>&1
fi
# the build trees are only intermediates
rm -rf "$R/mfem-build" "$R/axom-build"
echo "reference third-party libs ready under $R"
