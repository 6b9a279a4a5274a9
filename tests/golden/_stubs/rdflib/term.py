class URIRef(str):
    pass


class BNode(str):
    pass


class Literal(str):
    datatype = None
    language = None
