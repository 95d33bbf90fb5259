"""TEST INFRASTRUCTURE ONLY.

CPU oracle for the CenterMask2 inference path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import anything from here, and only
as the checker / the reported CPU baseline -- never on the product path.

* ``oracle.d2shim``   stand-in for detectron2 / fvcore / pycocotools so the reference's own files
                      run unchanged *in the build container* (``/root/reference`` is absent on the
                      GPU box).
* ``oracle.refrun``   builds and runs the unmodified reference model over the shim.
* ``oracle.restate``  independent fp32 restatement of the same path in plain torch functional ops;
                      travels to the GPU box; pinned against ``refrun`` by ``tests/golden``.

Parity pin: the reference ships no tests or golden vectors, so the pin is the
reference itself executed here: ``oracle/gen_golden.py`` writes ``tests/golden/*.pt``.
"""
