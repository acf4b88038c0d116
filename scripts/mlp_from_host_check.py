import sys, torch
sys.path.insert(0, ".")
import bench, torchode_b200 as to
w = bench.C4("c4", 8192)
host = {k: (None if v is None else v.pin_memory()) for k, v in w.host_inputs(0, 8192).items()}
field, method, ctrl = w.components("cuda")
solver = to.AutoDiffAdjoint(method, ctrl)
hp = to.InitialValueProblem(host["y0"], host["t_start"], host["t_end"], host["t_eval"])
with torch.no_grad():
    want = solver.solve(bench.make_problem(host, "cuda"))
    got = to.solve_from_host(solver, hp, "cuda")
    print("chunks", solver.last_run, "equal", torch.equal(got.ys, want.ys.cpu()), got.stats["n_accepted"].sum().item(), want.stats["n_accepted"].sum().item())
