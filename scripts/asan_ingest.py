import sys, os, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from besst_b200 import build
build.build_hostcheck = lambda force=False: "/tmp/libhc_asan.so"     # the sanitizer build instead of the in-tree one
import test_bamdev as t, test_bamio as tb, test_bamdev_parts as tp
from besst_b200 import bamio
import tempfile, pathlib
tmp = pathlib.Path(tempfile.mkdtemp())
for lvl in (0, 1, 6, 9): t.test_inflate_and_crc_equal_zlib(lvl)
t.test_inflate_fuzz_against_zlib()
t.test_inflate_rejects_damaged_streams()
for n, bb in [(0, 3000), (1, 3000), (5000, 700), (5000, 65000)]:
    t.test_host_rendered_ingest_equals_python_reader(tmp, n, bb)
t.test_host_rendered_ingest_errors(tmp)
for bb in (65280, 3000): t.test_host_rendered_ingest_of_records_longer_than_a_block(tmp, bb)
tp.test_parts_concatenate_to_the_whole_file(tmp, 5000, 700, 37)
tp.test_parts_concatenate_to_the_whole_file(tmp, 300, 700, 400)
tp.test_a_record_longer_than_the_tail_is_an_error(tmp)
print("ASAN/UBSAN ingest run complete")
