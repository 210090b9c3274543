class Figure:
    pass
