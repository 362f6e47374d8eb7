# bench_reference.jl — the TRUE reference number (Julia >= 1.12 + LevelSetMethods.jl v0.2.0 required; neither is in
# this repository's images, so bench.py reports the C++ oracle port instead and says so).
# Same workload as bench.py: C3 Enright sphere, WENO5 + RK3, NeumannBC, stored velocity x cos(pi t / 3), n^3 nodes.
#
#   JULIA_NUM_THREADS=$(nproc) julia --project baseline/bench_reference.jl 96 3
using LevelSetMethods, StaticArrays
const LSM = LevelSetMethods
n = parse(Int, get(ARGS, 1, "96")); steps = parse(Int, get(ARGS, 2, "3"))
g = CartesianGrid((0, 0, 0), (1, 1, 1), (n, n, n))
ϕ = MeshField(x -> sqrt(sum(abs2, x .- 0.35)) - 0.15, g)
u0 = MeshField(x -> SVector(2sin(π * x[1])^2 * sin(2π * x[2]) * sin(2π * x[3]), -sin(2π * x[1]) * sin(π * x[2])^2 * sin(2π * x[3]),
                            -sin(2π * x[1]) * sin(2π * x[2]) * sin(π * x[3])^2), g)
u = MeshField(copy(values(u0)), g)
upd = (vel, _ϕ, t) -> (values(vel) .= values(u0) .* cos(π * t / 3); nothing)     # the reference's way to make u time dependent
eq = LevelSetEquation(; terms = (AdvectionTerm(u, WENO5(), upd),), ic = ϕ, bc = NeumannBC(), integrator = RK3())
dt = 0.5 * LSM.compute_cfl(eq.terms, current_state(eq), 0.0)
integrate!(eq, dt)                                       # warm-up / compile
t0 = time(); integrate!(eq, current_time(eq) + steps * dt * 0.999); sec = time() - t0
println("JULIA_NUM_THREADS=", Threads.nthreads(), " CPU_THREADS=", Sys.CPU_THREADS,
        "  cell-updates/s = ", n^3 * steps / sec, "  (the hot loop in src/timestepping.jl:128-202 is serial)")
