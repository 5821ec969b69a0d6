#!/bin/bash
# AddressSanitizer + UBSan over the host-compiled cores (no GPU): the device ingest's source (bgzf_core.cuh, bam_ingest.hpp via
# bgzf_hostcheck.cpp) on the zlib / fuzz / damaged-stream / ragged-file / long-record / parts tests, besst_paths.cu (path search,
# scaffold pruning) on the reference-parity tests, and the host-thread BAM reader (besst_bamio.cpp) on tests/test_bamio.py.  Needs /root/reference for the second half.
set -e
cd "$(dirname "$0")/.."
FLAGS="-O1 -g -std=c++17 -shared -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer"
g++ $FLAGS -o /tmp/libhc_asan.so besst_b200/csrc/bgzf_hostcheck.cpp
g++ $FLAGS -x c++ -o /tmp/libpaths_asan.so besst_b200/csrc/besst_paths.cu
g++ $FLAGS -pthread -o /tmp/libbamio_asan.so besst_b200/csrc/besst_bamio.cpp -lz
export LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1
python scripts/asan_ingest.py
python scripts/asan_paths.py
python scripts/asan_bamio.py
# ThreadSanitizer over the threaded path search (8 threads, 1000 start nodes, the iteration cap hit)
unset LD_PRELOAD
g++ -O1 -g -std=c++17 -fsanitize=thread -pthread -x c++ besst_b200/csrc/besst_paths.cu scripts/tsan_paths.cpp -o /tmp/tsan_paths && /tmp/tsan_paths
