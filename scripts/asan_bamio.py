import sys, pathlib, tempfile
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from besst_b200 import bamio
bamio.BAMIO_SO = "/tmp/libbamio_asan.so"
import test_bamio as tb
tmp = pathlib.Path(tempfile.mkdtemp())
for n, bb, th in [(0, 3000, 2), (1, 3000, 1), (5000, 700, 4), (5000, 65000, 3), (20000, 3000, 8)]:
    tb.test_native_reader_equals_python_reader(tmp, n, bb, th)
import inspect
for name, fn in inspect.getmembers(tb, inspect.isfunction):
    if name.startswith("test_") and name not in ("test_native_reader_equals_python_reader",):
        params = list(inspect.signature(fn).parameters)
        if params in ([], ["tmp_path"]):
            try:
                fn(*([tmp] if params else []))
                print("ran", name)
            except BaseException as e:
                print("skip/err", name, type(e).__name__, str(e)[:80])
print("ASAN bamio run complete")
