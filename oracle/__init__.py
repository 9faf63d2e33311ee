"""CPU oracle for the HEOM/DEOM RK4 propagation path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pyqed_b200/`` may import this
package: it exists so that ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` have an independent
CPU statement of the reference algorithm to check the CUDA path against.

Parity status: PINNED.  The reference (binggu56/pyqed) ships no tests and no
golden vectors for this path, so the pin is against outputs
of the reference itself: ``tests/golden/make_golden.py`` imports the unmodified
reference from ``/root/reference`` (by file path) in the build container, runs
it on the inputs listed there and commits the results as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function of this oracle against
those fixtures.

Modules
-------
deom_oracle   multi-exponential DEOM (``pyqed/heom/deom.py:555-766, 1048-1114``)
heom_oracle.c the same DEOM RK4 path in plain C with an OpenMP loop over ADOs
              (``c_oracle.py`` builds and binds it); checker for mid-size cases
              and the multi-threaded CPU baseline of ``bench.py``
chain_oracle  single-exponential chain HEOM, RK4 (``pyqed/HEOM/heom.py:275-347``
              + ``pyqed/phys.py:1051-1064``) and Euler (``pyqed/oqs.py:1808-1875``)
"""
