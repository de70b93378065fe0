"""CPU oracle for the amaranth_twstft per-second PRN correlator.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (numpy fp64 + a few
lines of C for the LFSR) of the reference algorithm.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker or the timed CPU
baseline -- never as part of the shipped (CUDA) product path.  The product
package ``amaranth_twstft_b200`` never imports it.

Parity status: PINNED.
  * PRN generator: byte-exact against the 19 PRN files shipped in the reference
    (digests in tests/golden/prn_digests.json, made by tools/make_golden.py).
  * Parabolic fine delay: 15 050 (xvalm1,xval,xvalp1,correction) triples from
    experiments/220616_Besancon/*.mat.gz (tests/golden/parabola_kat.npz).
  * df grid quirk: df=843.7501687498 KAT from the same files.
  * End to end (variant G): outputs of the reference's own
    experiments/221219_twoway/processing/godual_ranging.py, executed unmodified
    in the authoring container on seeded synthetic records
    (tests/golden/ref_py_*.json, made by tools/make_golden.py).
All file:line citations are relative to the reference repository root.
"""
