"""CPU oracle for the ECoG preprocess -> epoch -> channel-selection hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``decode_tonal_langauge_b200/`` may
import this package; the only callers are ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, and there only as the checker or the timed CPU arm.

What it is: a numpy restatement of the reference's algorithms
(Daniel-Lin-S/decode_tonal_langauge, mounted read-only at /root/reference in the
build container), each function citing the reference ``file:line`` it follows.
The reference delegates most of its arithmetic to third-party wheels that are
NOT under /root/reference: ``scipy`` (pinned ``==1.11.4`` in the reference's
``requirements.txt:1``; this image has 1.18.1, which is the oracle of record),
``pandas`` (pinned 2.3.0, image 3.0.2) and numpy.  The scipy-level algorithms
(``filtfilt`` odd-extension + ``lfilter_zi``, ``resample``, ``f_oneway``) are
restated here from their published form; only scipy's *primitives* are used
(``lfilter`` = the C direct-form-II-transposed recursion, ``scipy.fft`` =
pocketfft, ``special.fdtrc``, and the ``butter``/``firwin`` designs).

Parity pinning: the reference has NO tests, golden vectors or known-answer
fixtures for this path.  The oracle is therefore
pinned against outputs of the reference itself, imported unmodified from
/root/reference in the build container (``oracle/reference_bridge.py``):
``tests/golden/make_golden.py`` wrote the committed ``tests/golden/*.npz``
fixtures from the real reference, ``tests/test_oracle_golden.py`` checks the
oracle against them everywhere, and ``tests/test_oracle_vs_reference.py``
checks it against the live reference when /root/reference is present.
"""

SCIPY_OF_RECORD = "1.18.1"
