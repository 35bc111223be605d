"""CPU oracle for the encoder + serialized-CTC hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the
product package (``multi-talker-asr-with-llms_b200/``).  Allowed importers:
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs.

Parity status: the reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF run in the build container: ``oracle/gen_golden.py`` imports the
reference's modules from /root/reference, runs them on seeded inputs, asserts
the restatements here agree, and writes ``tests/golden/*.npz``.
"""
