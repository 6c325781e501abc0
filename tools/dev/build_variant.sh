# usage: build_variant.sh <name> <-D flags...>   -> tools/_build/liborbgpu_<name>.so
name=$1; shift
cd "$(dirname "$0")/../../orb_slam2_with_comment_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -Xcompiler -fPIC,-O2,-ffp-contract=off -shared -cudart static "$@" -Xptxas -v \
  -o ../../tools/_build/liborbgpu_$name.so og_capi.cu og_match.cu og_vocab.cu og_multi.cu 2>&1 | grep -A2 "k_orient_desc\|k_resize4_pp" | grep "Used\|spill"
