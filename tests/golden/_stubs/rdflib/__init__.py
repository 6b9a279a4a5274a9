"""Import-only stand-in for rdflib, used solely by tests/golden/make_golden.py so that the
reference package (which imports rdflib at module scope, e.g. mrgcn/data/utils.py:10) can be
imported in a container without rdflib.  No RDF functionality is provided or needed."""


class Namespace(str):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return self + name

    def __getitem__(self, name):
        return self + str(name)
