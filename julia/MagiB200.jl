# MagiB200.jl -- the reference-side binding of libmagi_b200.so (see INTEGRATION.md).
# NOT executed in this repository's CI: Julia is not available in the build image.  The same C ABI is exercised through
# Python ctypes (tests/), entry point for entry point.
#
# Drop-in for the reference's MagiTarget (src/logdensityproblems_interface.jl:33-45): implements the four
# LogDensityProblems methods the NUTS loop of src/samplers.jl:137-138 calls, by `ccall` into the shared library.
module MagiB200

using LogDensityProblems

const LIB = get(ENV, "MAGI_B200_LIB", "libmagi_b200.so")

struct MagiConfig                      # mirrors magi_config in include/magi_b200.h (field order matters)
    n_times::Cint; n_dims::Cint; n_params_ode::Cint; kernel_id::Cint
    bandsize::Cint; ode_model_id::Cint; sigma_is_fixed::Cint; setup_mode::Cint
    max_chains::Cint; device::Cint
    jitter::Cdouble
    tvec::Ptr{Cdouble}; phi::Ptr{Cdouble}; yobs::Ptr{Cdouble}; sigma_init::Ptr{Cdouble}; prior_temperature::Ptr{Cdouble}
end

mutable struct MagiTargetGPU
    h::Ptr{Cvoid}
    P::Int
    function MagiTargetGPU(h::Ptr{Cvoid})
        t = new(h, Int(ccall((:magi_dimension, LIB), Cint, (Ptr{Cvoid},), h)))
        finalizer(x -> ccall((:magi_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), t)
        return t
    end
end

last_error() = unsafe_string(ccall((:magi_last_error, LIB), Cstring, ()))

const MODEL_IDS = Dict(:fn => 0, :hes1 => 1, :hes1log => 2, :hes1log_fixg => 3, :hes1log_fixf => 4, :hiv => 5, :ptrans => 6, :lv => 7, :lorenz96 => 8)

"""
    MagiTargetGPU(y_obs, t_obs, phi, model; sigma_init, prior_temperature, sigma_is_fixed, kernel, bandsize, jitter, ...)

Replaces steps 4-5 of `solve_magi` (src/MagiJl.jl:456-520): the per-dimension `calculate_gp_covariances!` calls and the
`MagiTarget(...)` construction.  `phi` is the 2 x D matrix of (variance; lengthscale) (src/MagiJl.jl:466-467).
"""
function MagiTargetGPU(y_obs::Matrix{Float64}, t_obs::Vector{Float64}, phi::Matrix{Float64}, model::Symbol;
                       sigma_init::Vector{Float64}, prior_temperature::Vector{Float64} = [1.0, 1.0, 1.0],
                       sigma_is_fixed::Bool = false, kernel::String = "matern52", bandsize::Int = 20, jitter::Float64 = 1e-6,
                       n_params_ode::Int, setup_mode::Int = 0, max_chains::Int = 1, device::Int = 0)
    n, D = size(y_obs)
    cfg = Ref(MagiConfig(n, D, n_params_ode, kernel == "rbf" ? 1 : 0, bandsize, MODEL_IDS[model], sigma_is_fixed ? 1 : 0,
                         setup_mode, max_chains, device, jitter, pointer(t_obs), pointer(phi), pointer(y_obs),
                         pointer(sigma_init), pointer(prior_temperature)))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = GC.@preserve y_obs t_obs phi sigma_init prior_temperature ccall((:magi_create, LIB), Cint, (Ref{MagiConfig}, Ref{Ptr{Cvoid}}), cfg, out)
    rc == 0 || error("magi_create failed: " * last_error())
    return MagiTargetGPU(out[])
end

LogDensityProblems.dimension(t::MagiTargetGPU) = t.P
LogDensityProblems.capabilities(::Type{MagiTargetGPU}) = LogDensityProblems.LogDensityOrder{1}()

function LogDensityProblems.logdensity(t::MagiTargetGPU, params::AbstractVector{Float64})
    ll = Ref{Cdouble}(0.0)
    p = params isa Vector{Float64} ? params : collect(params)
    rc = ccall((:magi_logdensity, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ref{Cdouble}), t.h, p, length(p), ll)
    rc == 0 || error("magi_logdensity failed: " * last_error())
    return ll[]
end

function LogDensityProblems.logdensity_and_gradient(t::MagiTargetGPU, params::AbstractVector{Float64})
    ll = Ref{Cdouble}(0.0)
    grad = Vector{Float64}(undef, t.P)
    p = params isa Vector{Float64} ? params : collect(params)
    rc = ccall((:magi_logdensity_and_gradient, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ref{Cdouble}, Ptr{Cdouble}), t.h, p, length(p), ll, grad)
    rc == 0 || error("magi_logdensity_and_gradient failed: " * last_error())
    return ll[], grad
end

"""
Batched evaluation: `params` is a P x n_chains Matrix (one chain per column, the vectorised-HMC shape of AdvancedHMC).
Returns (ll::Vector, grad::Matrix).
"""
function logdensity_and_gradient_batched(t::MagiTargetGPU, params::Matrix{Float64})
    @assert size(params, 1) == t.P
    nc = size(params, 2)
    ll = Vector{Float64}(undef, nc); grad = similar(params)
    rc = ccall((:magi_logdensity_and_gradient_batched, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), t.h, nc, params, ll, grad)
    rc == 0 || error("batched evaluation failed: " * last_error())
    return ll, grad
end

"""
    run_hmc_sampler(t, initial_params; n_samples, n_adapts, target_accept_ratio, initial_step_size, n_leapfrog, max_tree_depth, seed)

Many chains at once with the state resident on the GPU (`magi_hmc_*`): the counterpart of `run_nuts_sampler`
(src/samplers.jl:114-194) for a `P x n_chains` matrix of starting points.  `n_samples` is the total number of iterations
including the `n_adapts` warm-up iterations, which are dropped.  Returns `(draws, accept_rate, step_size)` with
`draws[:, chain, iteration]` = (theta..., sigma..., lp).  `max_tree_depth > 0` runs the reference's own trajectory
(multinomial NUTS with the generalised U-turn criterion, src/samplers.jl:158-160) batched on the device (`magi_nuts_run`) instead of
static trajectories of `n_leapfrog` steps.
"""
function run_hmc_sampler(t::MagiTargetGPU, initial_params::Matrix{Float64}; n_samples::Int = 2000, n_adapts::Int = 1000,
                         target_accept_ratio::Float64 = 0.8, initial_step_size::Float64 = 0.1, n_leapfrog::Int = 20,
                         max_tree_depth::Int = 0, seed::Integer = 0, chain_id_offset::Integer = 0, n_cols::Int)
    @assert size(initial_params, 1) == t.P "Initial parameters dimension mismatch"        # samplers.jl:125
    nc = size(initial_params, 2)
    chk(rc, what) = rc == 0 || error(what * " failed: " * last_error())
    chk(ccall((:magi_hmc_init, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Culonglong, Cdouble, Clonglong),
              t.h, nc, initial_params, seed, initial_step_size, chain_id_offset), "magi_hmc_init")
    # static trajectories: (n_iter, n_leapfrog, ...); NUTS: (n_iter, max_depth, ...) -- same argument list, two entry points
    run(n_iter, adapt, store) = max_tree_depth > 0 ?
        ccall((:magi_nuts_run, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cdouble, Cint, Ptr{Cvoid}), t.h, n_iter, max_tree_depth, adapt, target_accept_ratio, store, C_NULL) :
        ccall((:magi_hmc_run, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cdouble, Cint, Ptr{Cvoid}), t.h, n_iter, n_leapfrog, adapt, target_accept_ratio, store, C_NULL)
    n_adapts > 0 && chk(run(n_adapts, 1, 0), "sampler run (warm-up)")
    chk(ccall((:magi_hmc_reset_stats, LIB), Cint, (Ptr{Cvoid},), t.h), "magi_hmc_reset_stats")
    n_keep = n_samples - n_adapts
    n_keep > 0 && chk(run(n_keep, 0, 1), "sampler run")
    draws = Array{Float64}(undef, n_cols, nc, max(n_keep, 0))          # n_cols = n_params_ode + n_dims + 1
    stored = Ref{Clonglong}(0)
    chk(ccall((:magi_hmc_get_draws, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Clonglong, Ref{Clonglong}), t.h, draws, max(n_keep, 0), stored), "magi_hmc_get_draws")
    acc = Vector{Float64}(undef, nc); eps = Vector{Float64}(undef, nc); ndiv = Vector{Cint}(undef, nc)
    chk(ccall((:magi_hmc_get_stats, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}),
              t.h, acc, eps, ndiv, C_NULL, C_NULL), "magi_hmc_get_stats")
    return draws, acc, eps
end

"""
Multi-GPU runs (one Julia process per GPU, chains sharded over the processes): rank 0 draws the NCCL unique id with
`nccl_unique_id()`, the host distributes its 128 bytes (MPI.jl, a file, a socket), every rank calls `comm_init!`; the warm-up's
pooled statistics and `allgather_draws!` then run inside the library on NVLink.  `out` is a device pointer to
`world * n_stored * n_chains * n_cols` doubles (e.g. a CUDA.jl `CuArray`'s pointer; the library never touches CUDA.jl itself).
"""
function nccl_unique_id()
    id = Vector{UInt8}(undef, 128)
    rc = ccall((:magi_nccl_unique_id, LIB), Cint, (Ptr{UInt8},), id)
    rc == 0 || error("magi_nccl_unique_id failed: " * last_error())
    return id
end
function comm_init!(t::MagiTargetGPU, id::Vector{UInt8}, rank::Integer, world::Integer; n_chains_total::Integer)
    rc = ccall((:magi_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), t.h, id, rank, world)
    rc == 0 || error("magi_comm_init failed: " * last_error())
    ccall((:magi_comm_warmup, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), t.h, C_NULL)
    return n_chains_total          # pass it to magi_hmc_set_global after magi_hmc_init (run_hmc_sampler's chain_id_offset = first chain of this rank)
end
allgather_draws!(t::MagiTargetGPU, out::Ptr{Cdouble}) =
    ccall((:magi_hmc_allgather_draws, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cvoid}), t.h, out, C_NULL) == 0 || error(last_error())

"""`x_sampled` of solve_magi's result (src/MagiJl.jl:633-771): call `store_x!` after `magi_hmc_init` and before the kept iterations."""
store_x!(t::MagiTargetGPU, n_chains_x::Integer, thin::Integer = 1) =
    ccall((:magi_hmc_store_x, LIB), Cint, (Ptr{Cvoid}, Cint, Cint), t.h, n_chains_x, thin) == 0 || error(last_error())
function x_draws(t::MagiTargetGPU, n::Integer, D::Integer)
    ns = Ref{Clonglong}(0); ncx = Ref{Cint}(0)
    ccall((:magi_hmc_get_x_draws, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Clonglong, Ref{Clonglong}, Ref{Cint}), t.h, C_NULL, 0, ns, ncx)
    out = Array{Float64}(undef, n, D, Int(ncx[]), Int(ns[]))          # [time, dimension, chain, draw]
    ccall((:magi_hmc_get_x_draws, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Clonglong, Ref{Clonglong}, Ref{Cint}), t.h, out, ns[], ns, ncx)
    return out
end

end # module
