#!/bin/bash
# CMake build of the backend + facade + Catch2-style tests, then CTest, on the GPU box.
TAG=${1:-ctest}; O=gpurun_out/$TAG; mkdir -p $O
B=/tmp/blf_cmake_build
( cmake -S bipedal_locomotion_framework_b200/cpp -B $B -G Ninja -DBUILD_TESTING=ON && cmake --build $B -j 16 ) > $O/cmake_build.log 2>&1
echo "cmake exit $?" | tee -a $O/cmake_build.log
( cd $B && ctest --output-on-failure ) > $O/ctest.log 2>&1; echo "ctest exit $?" | tee -a $O/ctest.log
tail -12 $O/ctest.log
cuobjdump -lelf $B/libblf_ccm.so | head -3 | tee $O/cubin_arch.txt
