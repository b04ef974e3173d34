def get_one_exchange_neighbourhood(*a, **k):
    raise NotImplementedError('not needed by the offline (tabular) path exercised by the fixture generator')
