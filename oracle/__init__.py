"""CPU oracle: a numpy complex128 restatement of the reference's hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package (``tensornetworkevolve.jl_b200/``).  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
there only as the checker / the timed CPU baseline.

What it restates (file:line are relative to the reference checkout):

* ``src/peps.jl``            -> :mod:`oracle.peps`
* ``src/tebd.jl``            -> :mod:`oracle.tebd`
* ``src/timeevolve.jl``      -> :mod:`oracle.timeevolve`
* ``src/TensorAD/*.jl``      -> :mod:`oracle.tensorad`
* OMEinsum 0.7 (not vendored; ``Project.toml:19``): binary einsum as
  permutedims -> reshape -> batched gemm -> reshape -> permutedims, ``einsum_grad``,
  greedy contraction order, sliced execution -> :mod:`oracle.einsum`
* Yao 0.8 (not vendored; ``Project.toml:20``): the handful of blocks the path
  touches (put / kron / control / Scale / Add / time_evolve, little-endian dense
  register) -> :mod:`oracle.yao`
* Graphs 1.7 ``SimpleGraph`` (edges in lexicographic order, ascending neighbours)
  -> :mod:`oracle.graphs`

Parity pinning: the reference is Julia and cannot run in this image (no Julia runtime,
none of OMEinsum/Yao/KrylovKit vendored), and it stores no golden vectors.  Its own
tests are known-answer and dense-statevector identities; those are restated one by
one in ``tests/test_oracle_*.py`` (``test/peps.jl:11,17,19`` known answers, the
``statevec`` equalities, ``test/tebd.jl`` dense evolution, ``test/timeevolve.jl``
gradient convention / Hessian-vector product against an independent torch-autograd
restatement).  Parity is therefore pinned by identities against dense linear
algebra, not by stored outputs of a Julia run.
"""
