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
  * CAVIaR / simulate: PINNED against the reference's OWN SOURCE executed here -- the
    unmodified circuitmap/{model,simulation}.py and optimise/{caviar,pava}.py are imported
    from /root/reference through oracle/jax_shim.py (a NumPy-backed stand-in for the few
    JAX symbols they import; jax itself is not installable: no wheel, no network) by
    oracle/make_golden.py, and oracle/caviar.py / oracle/simulate.py are checked against
    the committed outputs (tests/golden/caviar_ref_*.npz, tests/test_reference_pin.py):
    per-iteration histories, every accept/reject decision, reconnect edge cases.
  * PRNG: the threefry stream under that run is the restatement in oracle/prng.py, pinned
    to the Random123 KATs and to the known answers of JAX's own test-suite
    (tests/test_oracle_prng.py) -- not to a live jax.random.  XLA's reduction order is
    likewise not reproducible here (the reference has no tests or golden vectors).
"""
