"""Developer probe: where does the host-buffer (e2e) step spend its time?  GMG_TRACE=1 python scripts/trace_e2e.py [n]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
class A: atoms_n = n
P = bench.pkg()
path, pos, q = bench.write_atoms(A)
os.environ.pop("GMG_TRACE", None)
B = P.hostapi.BenchProblem(P.lattice.cluster_prm(path, n, cycles=5))
B.step_host(True)
os.environ["GMG_TRACE"] = "1"
for name, fn in (("step_host(hierarchy)", lambda: B.step_host(True)), ("step_host(no hierarchy)", lambda: B.step_host(False)),
                 ("step_device", B.step_device)):
    t = time.time(); fn(); B.gmg.synchronize()
    print("== %s: %.1f ms" % (name, 1e3 * (time.time() - t)), file=sys.stderr)
# the same with the system / level-0 matrices assembled on the device (Matrix assembly = Device)
os.environ.pop("GMG_TRACE", None)
B.set_device_assembly(True)
B.step_host(True)
B.step_host(True)
os.environ["GMG_TRACE"] = "1"
t = time.time(); B.step_host(True); B.gmg.synchronize()
print("== step_host(hierarchy, device assembly): %.1f ms" % (1e3 * (time.time() - t)), file=sys.stderr)
