"""CPU oracle for the spiking-FireNet hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU (torch fp32 / numpy) restatement of the
reference algorithm (LSquarzoni/SNN_Event-based_Optical_Flow), each function
citing the reference file:line it follows.  It is the *checker*: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it.  The product package
(``snn_event-based_optical_flow_b200``) never imports it and has no CPU
fallback.

Parity pinning: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4), so the oracle is pinned against the reference itself:
``oracle/make_golden.py`` imports the unmodified reference from
``/root/reference`` (through the import shim in ``oracle/ref_shim.py``), runs it
on seeded synthetic inputs and writes ``tests/golden/*.npz``;
``tests/test_oracle_vs_golden.py`` checks the oracle against those fixtures
everywhere, and ``tests/test_oracle_vs_reference.py`` checks it against the live
reference wherever ``/root/reference`` exists.
"""
