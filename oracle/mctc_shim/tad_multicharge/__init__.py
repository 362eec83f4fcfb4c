"""
TEST INFRASTRUCTURE ONLY -- placeholder for ``tad-multicharge==0.5.0``
(``/root/reference/setup.cfg:37``).  EEQ charges are outside the hot path
(BASELINE.json north_star); every oracle / golden run passes ``q=`` explicitly.
"""


def get_eeq_charges(*args, **kwargs):
    raise NotImplementedError(
        "tad-multicharge is not available in this image; pass atomic charges "
        "explicitly via `q=`."
    )
