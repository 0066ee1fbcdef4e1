import os, sys, time
t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.chdir(ROOT)
out = open(os.path.join(ROOT, "gpurun_out", "r4a_restrict.log"), "w")
def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True); out.write(s + "\n"); out.flush()
try:
    import test_gpu_zz_restrict as T
    log("imports %.1fs" % (time.time() - t0))
    t = time.time(); T.test_restrict_reaction_candidates_are_bit_exact(); log("PASS test_restrict_reaction_candidates_are_bit_exact %.1fs" % (time.time() - t))
    import test_gpu_reactions as R
    t = time.time(); R.test_reaction_pass_p1_sets_are_bit_exact(0); log("PASS test_reaction_pass_p1_sets_are_bit_exact[0] %.1fs" % (time.time() - t))
    t = time.time(); R.test_acceptance_draws_match(); log("PASS test_acceptance_draws_match %.1fs" % (time.time() - t))
    import __graft_entry__ as G
    t = time.time(); G.smoke(); log("PASS smoke %.1fs" % (time.time() - t))
    import tempfile, pathlib
    t = time.time(); T.test_dacron_restrict_driver_gpu_matches_oracle(pathlib.Path(tempfile.mkdtemp())); log("PASS test_dacron_restrict_driver_gpu_matches_oracle %.1fs" % (time.time() - t))
    t = time.time(); T.test_mf_driver_gpu_matches_oracle(pathlib.Path(tempfile.mkdtemp())); log("PASS test_mf_driver_gpu_matches_oracle %.1fs" % (time.time() - t))
except BaseException as e:
    import traceback
    log("FAIL", repr(e)); log(traceback.format_exc())
log("total %.1fs" % (time.time() - t0))
