import sys, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
from besst_b200 import _lib
L = C.CDLL("/tmp/libpaths_asan.so")
vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
L.besst_paths_between.restype = vp
L.besst_paths_between.argtypes = [i64, vp, vp, vp, vp, vp, i64, i64, C.c_double, i32, i32, i32]
L.besst_paths_count.restype = i64; L.besst_paths_count.argtypes = [vp]
L.besst_paths_hit_threshold.argtypes = [vp]
L.besst_paths_pops.restype = i64; L.besst_paths_pops.argtypes = [vp]
L.besst_paths_arrays.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
L.besst_paths_free.argtypes = [vp]
L.besst_scaffold_prune_ambiguous.restype = i64
L.besst_scaffold_prune_ambiguous.argtypes = [i64, i64, vp, vp, vp, vp, vp, vp, vp]
_lib._lib = L     # the sanitizer build behind besst_b200.ExtendLargeScaffolds / MakeScaffolds
import test_path_search as TP, test_scaffold_passes as TS
from besst_b200 import synth
import ref_harness
ref_harness.load_reference()
graphs = {}
for name, (n_contigs, n_pairs, seed) in {"mp_a": (600, 150000, 21), "mp_b": (1500, 300000, 22)}.items():
    lib = synth.make_library(n_contigs, n_pairs, "rf", 3000.0, 500.0, 0.0, seed=seed)
    r = ref_harness.run_reference(lib.to_batch(), dict(orientation="rf", mean=3000.0, stddev=500.0, readlen=100))
    graphs[name] = (r["objects"]["G_prime"], set(r["objects"]["Scaffolds"]))
TP.test_iteration_cap_and_ties(graphs)
for name in ("mp_a", "mp_b"):
    TP.test_between_scaffolds_equals_reference_on_pe_graphs(graphs, name)
TP.test_library_entry_point_without_reference_objects()
for trial in range(8):
    TS.test_passes_on_graphs_made_to_be_hard(trial)
print("ASAN/UBSAN paths + passes run complete")
