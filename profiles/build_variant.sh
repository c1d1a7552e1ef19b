#!/bin/bash
# build a tuning variant of libsrfuse next to the product library:  profiles/build_variant.sh <tag> <extra nvcc flags...>
#   -> super-resolution-system_b200/libsrfuse_<tag>.so ; select it at run time with SRF_LIB=<path> (see _native.lib_path)
set -e
tag=$1; shift
cd "$(dirname "$0")/../super-resolution-system_b200"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --threads 7 -shared "$@" \
     -o libsrfuse_${tag}.so csrc/srf_api.cu csrc/srf_fuse.cu csrc/srf_iqa.cu csrc/srf_pipeline.cu csrc/srf_resize.cu csrc/srf_merge.cu csrc/srf_seam.cu csrc/srf_comm.cu
echo "built libsrfuse_${tag}.so"
