# oracle/stubs/DBI.pm -- TEST INFRASTRUCTURE ONLY.
# Enough of DBI for SeqToolBox::Taxonomy->new and the gi_desc lookup of bin/ppp.pl to run
# ("no rows"); DBI/DBD::SQLite are not installed in this image.
package DBI;
sub connect { bless {}, 'DBI::stubdb' }
package DBI::stubdb;
sub prepare    { bless {}, 'DBI::stubst' }
sub disconnect { 1 }
sub DESTROY    { }
package DBI::stubst;
sub execute        { 1 }
sub fetchrow_array { () }
sub finish         { 1 }
sub DESTROY        { }
1;
