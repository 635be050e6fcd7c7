#!/bin/bash
# build a scratch copy of the tree under /tmp/dev (compile checks while a gpurun call is pending: the in-tree .so
# travels with the snapshot and must not be rewritten under it)
mkdir -p /tmp/dev
cd /root/repo && tar --exclude=.git --exclude=gpurun_out --exclude='*.so' --exclude='_obj' -cf - . | (cd /tmp/dev && tar -xf -)
cd /tmp/dev && python -m atomsmm_b200.build "$@" 2>&1 | tail -5
