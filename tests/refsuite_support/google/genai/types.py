class _Bag:
    def __init__(self, *args, **kwargs):
        self.__dict__.update(kwargs)


class GenerateContentConfig(_Bag):
    pass


class HttpOptions(_Bag):
    pass


class Part(_Bag):
    @classmethod
    def from_uri(cls, **kwargs):
        return cls(**kwargs)

    @classmethod
    def from_text(cls, **kwargs):
        return cls(**kwargs)


class Content(_Bag):
    pass
