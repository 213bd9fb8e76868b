class Dict(dict):
    """Attribute-style dict (codae/tool/dictionnary.py:5-9 of the reference)."""
    __getattr__ = dict.get
    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__
