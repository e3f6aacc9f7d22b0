"""CPU oracle for circuitmap_b200 -- TEST INFRASTRUCTURE, not product code.

NumPy fp64 restatement of the reference's CAVIaR path (circuitmap/optimise/caviar.py,
pava.py, model.py), of the JAX threefry PRNG it consumes, of circuitmap/simulation.py
(data generator) and of the NWD U-Net forward (circuitmap/neural_waveform_demixing.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package, and only as the checker.  The product (circuitmap_b200) never
imports it and has no CPU fallback.

Parity status
  * NWD: PINNED -- oracle/nwd.py is checked against the unmodified reference module
    (imported through a pytorch_lightning shim, oracle/make_golden.py) and against
    golden vectors committed under tests/golden/.
  * CAVIaR / simulate / PRNG: PARITY UNPINNED against true JAX -- jax is not installed
    (and not installable: no wheel, no network) and the reference holds no tests or
    golden vectors.  The threefry block function is pinned by the Random123 KATs and
    by values printed in JAX's public docs; everything else is a restatement that
    follows the cited reference lines.
"""
