"""ORACLE ONLY: empty stand-in (reference: snps_get_root_go_by_html.py:2, an offline HTML scraper off the hot path)."""


class BeautifulSoup:  # never instantiated
    pass
