#!/bin/bash
# Round 2, fourth session: FloatingBaseDynamicalSystem facade on the GPU.
TAG=${1:-r4c}
O=gpurun_out/$TAG
mkdir -p $O
./bipedal_locomotion_framework_b200/lib/FloatingBaseSystemDynamicsUnitTests > $O/cpp_FloatingBaseSystemDynamicsUnitTests.log 2>&1; echo "exit $?"
tail -25 $O/cpp_FloatingBaseSystemDynamicsUnitTests.log
timeout 600 python -m pytest tests/test_cpp_facade.py -m gpu -q -x 2>&1 | tail -5
