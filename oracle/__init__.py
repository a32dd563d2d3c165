"""oracle -- CPU restatement of momlevel's steric sea-level path (TEST INFRASTRUCTURE).

This package is the *checker*, never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  Nothing under ``momlevel_b200/`` imports it and the
product path raises when the CUDA library is missing rather than falling back here.

What it restates (reference = jkrasting/momlevel, paths relative to its root):

* ``oracle.eos``      -- ``src/momlevel/eos/wright.py:6-165``, ``src/momlevel/eos/linear.py:17-162``
* ``oracle.spice``    -- ``src/momlevel/spice/flament.py:7-95``
* ``oracle.steric``   -- ``src/momlevel/steric.py:84-184``, ``src/momlevel/reference.py:48-85``,
  ``src/momlevel/derived.py:249-325,414-444,642-666,769-795`` with xarray's
  name-based broadcasting / ``skipna`` sums written out in plain numpy
* ``oracle.stratification`` -- ``src/momlevel/derived.py:30-71, 328-411`` (cell-centre N2, "next" row)
* ``oracle.testdata`` -- ``src/momlevel/test_data/__init__.py:16-140``,
  ``test_data/tripolar/horizontal.py:110-115``, ``test_data/tripolar/vertical.py:37-68``

Parity pin status: PINNED.  ``tests/test_oracle.py`` checks the restatement against
(a) the reference's own known-answer values (``tests/test_wright.py``,
``tests/test_linear.py``, ``tests/test_flament.py``, ``tests/test_steric.py`` local
sums and reference-state sums, ``tests/test_derived.py`` dz / spice / masso sums) and
(b) ``tests/golden/*.npz``, input/output vectors produced by importing the reference's
numpy modules in the build container (``tests/golden/make_golden.py``).  The
reference's *global-domain* constants are below its own ``atol`` and therefore pin
nothing; for those the oracle follows the reference code and
the golden file records what that code evaluates to.
"""

from . import eos, spice, steric, stratification, testdata  # noqa: F401
