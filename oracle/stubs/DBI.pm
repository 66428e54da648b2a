# oracle/stubs/DBI.pm -- TEST INFRASTRUCTURE ONLY.
# Enough of DBI for SeqToolBox::Taxonomy->new and the gi_desc lookup of bin/ppp.pl to run ("no rows"); DBI/DBD::SQLite
# are not installed in this image.  If $ENV{PPP_STUB_GI2TAXID} names a two-column TSV (gi, tax_id), every statement
# answers execute($gi) with that one row -- what SeqToolBox::Taxonomy::get_taxon (lib/SeqToolBox/Taxonomy.pm:362-405)
# reads from gi_taxid_prot -- so that bin/ppp_cross_score.pl can be run on synthetic BLAST tables.  If $ENV{PPP_STUB_NODES}
# names a three-column TSV (tax_id, parent_tax_id, rank), statements whose SQL reads `from nodes` answer execute($tax_id) with
# that (parent, rank) row -- what SeqToolBox::Taxonomy::collapse_taxon (Taxonomy.pm:296-360) walks -- so that `ppp.pl -l <rank>`
# can be run unmodified.
package DBI;
my ( %map, %nodes );
if ( $ENV{PPP_STUB_NODES} && open( my $nh, '<', $ENV{PPP_STUB_NODES} ) ) {
	while (<$nh>) { chomp; my ( $t, $par, $rank ) = split /\t/; $nodes{$t} = [ $par, $rank ] if defined $rank; }
	close $nh;
}
if ( $ENV{PPP_STUB_GI2TAXID} && open( my $fh, '<', $ENV{PPP_STUB_GI2TAXID} ) ) {
	while (<$fh>) { chomp; my ( $g, $t ) = split /\t/; $map{$g} = $t if defined $t; }
	close $fh;
}
sub connect { bless {}, 'DBI::stubdb' }
package DBI::stubdb;
sub prepare    { my ( $d, $sql ) = @_; bless { row => undef, nodes => ( defined $sql && $sql =~ /from\s+nodes/i ) ? 1 : 0 }, 'DBI::stubst' }
sub disconnect { 1 }
sub DESTROY    { }
package DBI::stubst;
sub execute {
	my ( $s, $id ) = @_;
	if ( $s->{nodes} ) { $s->{row} = ( defined $id && exists $nodes{$id} ) ? $nodes{$id} : undef; return 1; }
	$s->{row} = ( defined $id && exists $map{$id} ) ? [ $map{$id} ] : undef;
	1;
}
sub fetchrow_array { my $s = shift; my $r = $s->{row}; $s->{row} = undef; return $r ? @$r : (); }
sub finish         { 1 }
sub DESTROY        { }
1;
