#!/bin/bash
# build tuning variants of the engine under /tmp/dev and collect them in atomsmm_b200/variants/ (git-ignored .so files that
# travel with the gpurun snapshot):   scripts/build_variants.sh name "-DFLAG=.. -DFLAG=.." [name flags ...]
mkdir -p /tmp/dev /root/repo/atomsmm_b200/variants
tar --exclude=.git --exclude=gpurun_out --exclude='*.so' --exclude='_obj' -cf - . | (cd /tmp/dev && tar -xf -)
while [ $# -ge 2 ]; do
    name="$1"; flags="$2"; shift 2
    rm -rf /tmp/dev/atomsmm_b200/csrc/_obj
    (cd /tmp/dev && B2_EXTRA_NVCC_FLAGS="$flags" B2_BUILD_OUTPUT=/root/repo/atomsmm_b200/variants/lib_$name.so python -m atomsmm_b200.build >/dev/null) || exit 1
    echo "== $name ($flags)"; grep -A2 "k_pair_force2I9LJCForce2ILi2ELi0ELi0ELi0ELi0\|k_pair_force2I9LJCForce2ILi1ELi0ELi1ELi0ELi2\|k_build_lists" /tmp/dev/atomsmm_b200/csrc/_obj/pair.o.log /tmp/dev/atomsmm_b200/csrc/_obj/nlist.o.log | grep "Used\|spill" | sed 's/.*ptxas info *: //; s/.*://'
done
