"""CPU oracle for the NeuroAlpha decoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker or the CPU arm that is timed *beside* the CUDA path.  The product
package (``neural_speech_decoding_b200``) never imports this package and
raises if its CUDA library is missing.

Parity status: PINNED.  The restatements here are checked (tests/test_oracle.py)
against golden vectors produced by running the *real* reference
(``/root/reference/Neuro-Alpha-App/Utilities/lstm_eeg_model.py`` and
``Utilities/tester.py``, imported unmodified by ``oracle/make_golden.py``) on the
reference's own checkpoint and its 324 ``EEG_data_collection`` windows.  The
reference has no tests of its own, so those generated vectors
are the pin.
"""
from .numpy_oracle import (  # noqa: F401
    lstm_layer_forward,
    decoder_forward,
    head_forward,
    softmax,
    trial_mean,
    zscore_window,
    RRELU_EVAL_SLOPE,
)
