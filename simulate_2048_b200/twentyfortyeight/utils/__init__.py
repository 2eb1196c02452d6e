"""Observation / reward presentation helpers — mirrors ``twentyfortyeight/utils`` of the reference
(``binary.py``, ``normalize.py``; the Matplotlib window of ``windows.py`` is out of scope)."""
