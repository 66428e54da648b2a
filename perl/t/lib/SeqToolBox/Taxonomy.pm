package SeqToolBox::Taxonomy;

# perl/t/lib/SeqToolBox/Taxonomy.pm -- TEST STUB (never installed): stands in for the reference's SeqToolBox::Taxonomy,
# which needs DBI + DBD::SQLite and the NCBI taxonomy database.  collapse_taxon walks the (tax_id, parent, rank) rows of the
# TSV named by $ENV{PPP_MOCK_NODES} the way lib/SeqToolBox/Taxonomy.pm:296-360 walks its nodes table; get_taxon answers
# gi -> tax_id from the two-column TSV named by $ENV{PPP_MOCK_GI2TAXID} (Taxonomy.pm:362-405: "gi|N|..." or a bare number).
use strict;
use warnings;

sub new {
    my $class = shift;
    my %rows;
    if ( $ENV{PPP_MOCK_NODES} && open( my $fh, '<', $ENV{PPP_MOCK_NODES} ) ) {
        while (<$fh>) { chomp; my ( $t, $p, $r ) = split /\t/; $rows{$t} = [ $p, $r ] if defined $r; }
        close $fh;
    }
    my %gi;
    if ( $ENV{PPP_MOCK_GI2TAXID} && open( my $fh, '<', $ENV{PPP_MOCK_GI2TAXID} ) ) {
        while (<$fh>) { chomp; my ( $g, $t ) = split /\t/; $gi{$g} = $t if defined $t; }
        close $fh;
    }
    return bless { rows => \%rows, gi => \%gi }, $class;
}

sub get_taxon {
    my ( $self, $gi ) = @_;
    die "Gi required in get_taxon\n" unless $gi;
    my $id;
    if    ( $gi =~ /gi\|(\d+)/i ) { $id = $1 }
    elsif ( $gi =~ /^(\d+)$/ )    { $id = $gi }
    else                          { die "Malformed GI: $gi\n" }
    return $self->{gi}{$id};
}

sub collapse_taxon {
    my ( $self, $id, $rank ) = @_;
    die "ERROR: ID not defined"   unless $id;
    die "ERROR: RANK not defined" unless $rank;
    my $taxid = $id;
    while ( $taxid && $taxid ne "1" ) {
        my $row = $self->{rows}{$taxid} or return undef;
        return $taxid if lc( $row->[1] ) eq lc($rank);
        $taxid = $row->[0];
    }
    return undef;
}

1;
