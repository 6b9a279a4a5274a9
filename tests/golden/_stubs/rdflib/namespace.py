from . import Namespace

XSD = Namespace("http://www.w3.org/2001/XMLSchema#")
