"""ConfigSpace.util stand-in: the lazy one-exchange neighbourhood generator over the restated stream
(oracle/ei_optimization.py ``OracleSpace.neighbourhood``); same keyword names and defaults as ConfigSpace's."""


def get_one_exchange_neighbourhood(configuration, seed, num_neighbors=4, stdev=0.2):
    from ConfigSpace import Configuration
    cs = configuration.configuration_space
    for row in cs._oracle_space().neighbourhood(configuration.get_array(), seed, num_neighbors=num_neighbors,
                                                stdev=stdev):
        yield Configuration(cs, vector=row)
