class Graph:
    pass
