"""CPU oracle for the Dia decode path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline - never as the thing shipped or measured as the GPU path.

Modules
-------
``dia_oracle``   functional torch-CPU fp32 restatement of the reference's
                 encoder / decoder / decode loop (dia/layers.py, dia/state.py,
                 dia/model.py).  Pinned against the patched reference itself,
                 see ``oracle/validate_against_reference.py``.
``delay_oracle`` numpy restatement of the integer delay / revert gathers
                 (dia/audio.py), pinned against the unpatched reference and the
                 known-answer vector in tests/golden/.
``delay_ref.c``  plain C restatement of the same integer path (built by
                 ``oracle/Makefile`` into ``oracle/_build/libdelay_ref.so``).
"""
