"""ncu driver: main-loss shape only (see ce_main_prof.py)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ce_main_prof as m
a, b, bias, tgt, col_ids, own = m.problem(103976, 21435)
args = (10.0, bias, tgt, col_ids, None, None, 0, float("-inf"), m.NO_DIAG)
wl = torch.full((a.shape[0],), 1.0 / a.shape[0], device="cuda")
out = torch.ops.rs.ce_fwd(a, b, *args, m.BOUND)
g = torch.ops.rs.ce_bwd(a, b, *args, out[0], wl, None, None, m.BOUND)
torch.cuda.synchronize()
