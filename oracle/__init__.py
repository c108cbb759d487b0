"""TEST INFRASTRUCTURE ONLY — CPU restatements ("oracles") of the reference hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.  The product
package ``tiatoolbox_b200`` never imports this package and fails loudly when its CUDA
library is missing.
"""
